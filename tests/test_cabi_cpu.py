"""CPU tests: the C-ABI library builds, loads and exports every symbol include/pb2.h declares; the
host-side key algebra (pure C inside libpb2) matches the oracle; host logic of the Python mirror.
No compute entry point is called (no GPU here)."""
import ctypes as C
import collections

import numpy as np
import pytest

from oracle import rng as orng


@pytest.fixture(scope='module')
def lib(lib_built):
  from probability_b200 import _lib
  return _lib.load()


def test_library_exports_every_declared_symbol(lib):
  from probability_b200 import _lib
  names = _lib.exported_symbols()
  assert len(names) >= 24
  for n in names:
    assert hasattr(lib, n), 'libpb2.so does not export ' + n
  assert lib.pb2_version() == 100


def test_ctx_create_fails_loudly_without_gpu(lib):
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  h = C.c_void_p()
  rc = lib.pb2_ctx_create(0, C.byref(h))
  assert rc != 0
  assert b'CUDA' in lib.pb2_last_error(None) or b'device' in lib.pb2_last_error(None)
  import probability_b200 as tfp
  from probability_b200 import _lib
  with pytest.raises(_lib.Pb2Error):
    tfp.targets.EightSchools()(0.0, 0.0, np.zeros(8, np.float32))
  with pytest.raises(_lib.Pb2Error):
    tfp.mcmc.effective_sample_size(np.zeros((10, 2), np.float32))


@pytest.mark.parametrize('layout', [0, 1])
def test_host_split_and_fold_in_match_oracle(lib, layout):
  import probability_b200 as tfp
  for seed in (0, 17, 123456789012):
    k = orng.key(seed)
    for n in (1, 2, 3, 4, 7):
      np.testing.assert_array_equal(tfp.random.split_seed(k, n=n, layout=layout), orng.split(k, n, layout))
    for d in (0, 1, 1365385517, 2**32 - 1):
      np.testing.assert_array_equal(tfp.random.fold_in(k, d), orng.fold_in(k, d))
  np.testing.assert_array_equal(tfp.random.sanitize_seed(17, salt='mcmc.sample_chain'), [2648418349, 421598061])
  with pytest.raises(TypeError):
    tfp.random.sanitize_seed(17, salt=3)
  with pytest.raises(TypeError):
    tfp.random.split_seed(17, n=2.0)


def test_python_callable_and_bad_args_raise(lib):
  import probability_b200 as tfp
  with pytest.raises(TypeError):
    tfp.mcmc.NoUTurnSampler(lambda x: x, step_size=1.)
  with pytest.raises(ValueError):
    tfp.mcmc.NoUTurnSampler(tfp.targets.EightSchools(), step_size=1., max_tree_depth=0)
  with pytest.raises(ValueError):
    tfp.mcmc.HamiltonianMonteCarlo(tfp.targets.EightSchools(), step_size=1., num_leapfrog_steps=0)
  k = tfp.mcmc.NoUTurnSampler(tfp.targets.EightSchools(), step_size=1., max_tree_depth=4)
  np.testing.assert_array_equal(k.write_instruction, [0, 4, 1, 4, 1, 4, 2, 4, 1, 4, 2, 4, 2, 4, 3, 4])
  k2 = k.copy(step_size=0.5)
  assert k2.step_size == 0.5 and k2.max_tree_depth == 4 and k.step_size == 1.


def test_targets_host_side(lib):
  import probability_b200 as tfp
  from oracle import targets as otargets
  t = tfp.targets.IllConditionedGaussian()
  cov, ev = otargets.ill_conditioned_covariance(100)
  np.testing.assert_allclose(t.covariance, cov)
  P, c = otargets.gaussian_precision_from_cov(cov)
  np.testing.assert_array_equal(t.precision, P)
  np.testing.assert_allclose(t.log_normalizer, c, rtol=1e-6)
  X, y = tfp.targets.synthetic_logistic_data(50, 4, seed=0)
  Xo, yo = otargets.synthetic_logistic_data(50, 4, seed=0)
  lr = tfp.targets.LogisticRegression(X, y)
  np.testing.assert_array_equal(lr.features_with_bias, Xo)
  np.testing.assert_array_equal(lr.labels, yo)
  np.testing.assert_array_equal(tfp.targets.synthetic_sv_returns(64), otargets.synthetic_sv_returns(64))
  assert tfp.targets.EightSchools().part_sizes == [1, 1, 8]
  assert tfp.targets.StochasticVolatility(np.zeros(10, np.float32)).dim == 13


# ---- sample_chain driver semantics with a kernel that needs no GPU (sample_test.py:57-99) ----
def test_sample_chain_loop_semantics_cpu(lib):
  import torch
  import probability_b200 as tfp
  Res = collections.namedtuple('Res', 'counter_1, counter_2, seed')

  class CountingKernel(tfp.mcmc.TransitionKernel):
    """state + 1 per step (sample_test.py TestTransitionKernel)."""
    is_calibrated = True

    def one_step(self, state, pkr, seed=None):
      return state + 1, Res(pkr.counter_1 + 1, pkr.counter_2 + 2, seed)

    def bootstrap_results(self, state):
      return Res(torch.tensor(0), torch.tensor(0), np.zeros(2, np.uint32))

  k = CountingKernel()
  st, tr = tfp.mcmc.sample_chain(5, torch.tensor(0.), kernel=k, num_burnin_steps=2, num_steps_between_results=1,
                                 trace_fn=lambda s, r: (r.counter_1, r.seed), seed=17)
  np.testing.assert_array_equal(st.numpy(), [3, 5, 7, 9, 11])      # 1+burnin, then 1+thin per result
  np.testing.assert_array_equal(tr[0].numpy(), [3, 5, 7, 9, 11])
  # seeds seen by one_step follow split(sanitize(seed, 'mcmc.sample_chain')) (sample.py:312,344-349)
  seed = orng.sanitize_seed(17, salt='mcmc.sample_chain')
  expect = []
  for i in range(11):
    s, seed = orng.split(seed, 2)
    expect.append(s)
  np.testing.assert_array_equal(tr[1], np.stack([expect[2], expect[4], expect[6], expect[8], expect[10]]))
  # no trace -> states only; checkpointable variant returns final results (sample_test.py:271-291)
  only = tfp.mcmc.sample_chain(3, torch.tensor(0.), kernel=k, trace_fn=None, seed=1)
  np.testing.assert_array_equal(only.numpy(), [1, 2, 3])
  res = tfp.mcmc.sample_chain(3, torch.tensor(0.), kernel=k, trace_fn=None, seed=1, return_final_kernel_results=True)
  assert int(res.final_kernel_results.counter_2) == 6
  res2 = tfp.mcmc.sample_chain(2, res.all_states[-1], previous_kernel_results=res.final_kernel_results, kernel=k,
                               trace_fn=None, seed=2, return_final_kernel_results=True)
  np.testing.assert_array_equal(res2.all_states.numpy(), [4, 5])
  assert int(res2.final_kernel_results.counter_1) == 5


def test_trace_fn_probe(lib):
  from probability_b200.mcmc import sample as s
  out = s._probe_trace_fn(lambda st, kr: {'a': kr.is_accepted, 'b': (kr.inner_results.step_size,)}, None)
  assert out is not None
  struct, paths = out
  assert paths == [('is_accepted',), ('inner_results', 'step_size')]
  assert s._probe_trace_fn(lambda st, kr: kr.log_accept_ratio + 1.0, None) is None
  assert s._probe_trace_fn(lambda st, kr: kr, None)[1] == [()]
