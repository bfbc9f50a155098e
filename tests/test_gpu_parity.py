"""GPU parity tests: every call goes through the C ABI (ctypes -> libpb2.so) and is compared
with the CPU oracle on the same seeded inputs.  Tolerances: uint32 RNG streams bit-exact;
single leapfrog trajectories 1e-5 relative (north_star); transition decisions identical
except where float rounding flips a comparison (bounded fraction, stated per test)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu

from oracle import diagnostic as odiag  # noqa: E402
from oracle import mcmc as omcmc  # noqa: E402
from oracle import rng as orng  # noqa: E402
from oracle import targets as otargets  # noqa: E402


@pytest.fixture(scope='module')
def tfp():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  import probability_b200 as tfp_
  return tfp_


def dev():
  return torch.device('cuda', 0)


def t(a):
  return torch.tensor(np.asarray(a), device=dev())


def rel_err(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return np.max(np.abs(a - b) / (np.abs(b) + 1e-3 * np.max(np.abs(b)) + 1e-30))


# ------------------------------------------------------------------ RNG
@pytest.mark.parametrize('layout', [0, 1, 2])   # threefry partitionable / original, Philox4x32-10
@pytest.mark.parametrize('n', [1, 2, 7, 64, 1001, 65536])
def test_rng_streams_bit_exact(tfp, layout, n):
  for seed in (0, 17, 2**40 + 5):
    k = orng.key(seed)
    got = tfp.random.bits(k, (n,), device=dev(), layout=layout).cpu().numpy()
    np.testing.assert_array_equal(got, orng.bits(k, n, layout))
    u = tfp.random.uniform((n,), seed=k, device=dev(), layout=layout).cpu().numpy()
    np.testing.assert_array_equal(u, orng.uniform(k, (n,), layout=layout))
    u2 = tfp.random.uniform((n,), minval=-2.0, maxval=3.0, seed=k, device=dev(), layout=layout).cpu().numpy()
    np.testing.assert_array_equal(u2, orng.uniform(k, (n,), -2.0, 3.0, layout=layout))
    z = tfp.random.normal((n,), seed=k, device=dev(), layout=layout).cpu().numpy()
    np.testing.assert_allclose(z, orng.normal(k, (n,), layout), rtol=2e-6, atol=2e-7)
    r = tfp.random.uniform((n,), 0, 2, dtype=torch.int32, seed=k, device=dev(), layout=layout).cpu().numpy()
    np.testing.assert_array_equal(r, orng.randint_bit(k, n, layout))
    r5 = tfp.random.uniform((n,), -3, 11, dtype=torch.int32, seed=k, device=dev(), layout=layout).cpu().numpy()
    np.testing.assert_array_equal(r5, orng.randint(k, n, -3, 11, layout))


def test_rng_known_answers(tfp):
  # jax-documented values
  z = tfp.random.normal((1,), seed=orng.key(0), device=dev(), layout=1).cpu().numpy()
  np.testing.assert_allclose(z, [-0.20584227], rtol=1e-6)
  np.testing.assert_array_equal(tfp.random.split_seed(orng.key(0), layout=1),
                                [[4146024105, 967050713], [2718843009, 1272950319]])
  np.testing.assert_array_equal(tfp.random.split_seed(orng.key(0), layout=0),
                                [[1797259609, 2579123966], [928981903, 3453687069]])
  # the JAX outputs the reference's executed notebook holds (tests/golden/jax_notebook_rng.json, make_golden.py):
  # the CUDA generator reproduces split / normal on the key and on both children
  import json
  import os
  g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'jax_notebook_rng.json')))
  kids = tfp.random.split_seed(orng.key(0), layout=1)
  np.testing.assert_array_equal(kids, g['split_key_0'])
  draw = lambda kk: float(tfp.random.normal((1,), seed=kk, device=dev(), layout=1).cpu().numpy()[0])
  np.testing.assert_allclose(draw(orng.key(0)), g['normal_key_0'], rtol=2e-7)
  np.testing.assert_allclose([draw(kids[0]), draw(kids[1])], g['normal_split_keys'], rtol=2e-7)
  m = g['more']
  u = tfp.random.uniform((1, 2), seed=orng.key(0), device=dev(), layout=1).cpu().numpy()
  np.testing.assert_allclose(u, [m['uniform_1x2_key_0']['value']], rtol=2e-7)
  z = tfp.random.normal((2, 5), seed=orng.key(0), device=dev(), layout=1).cpu().numpy()
  np.testing.assert_allclose(np.exp(z), m['exp_normal_2x5_key_0']['value'], rtol=1e-6)
  z8 = tfp.random.normal((8,), seed=orng.key(0), device=dev(), layout=1).cpu().numpy()
  np.testing.assert_allclose(z8, m['normal_8_key_0_tpu']['value'], rtol=2e-5)      # TPU erfinv: last digits


def test_rng_2d_shape_row_major(tfp):
  k = orng.key(5)
  got = tfp.random.normal((13, 7), seed=k, device=dev()).cpu().numpy()
  np.testing.assert_allclose(got, orng.normal(k, (13, 7)), rtol=2e-6, atol=2e-7)


# ------------------------------------------------------------------ targets
def _targets(tfp, which, seed=0):
  rng = np.random.default_rng(seed)
  if which == 'eight_schools':
    return tfp.targets.EightSchools(), otargets.EightSchools(), otargets.EightSchools(dtype=np.float64), \
        (np.array([0, 0] + [1] * 8) + 0.5 * rng.standard_normal((257, 10))).astype(np.float32)
  if which == 'dense3':
    cov = np.array([[1.0, 0.5, 0.1], [0.5, 2.0, -0.3], [0.1, -0.3, 0.5]])
    loc = np.array([1.0, -2.0, 0.5])
    tg = tfp.targets.DenseGaussian(covariance=cov, loc=loc)
    return tg, otargets.DenseGaussian(tg.precision, tg.log_normalizer, loc), \
        otargets.DenseGaussian(tg.precision.astype(np.float64), tg.log_normalizer, loc, dtype=np.float64), \
        rng.standard_normal((100, 3)).astype(np.float32)
  if which == 'dense100':
    tg = tfp.targets.IllConditionedGaussian()
    return tg, otargets.DenseGaussian(tg.precision, tg.log_normalizer), \
        otargets.DenseGaussian(tg.precision.astype(np.float64), tg.log_normalizer, dtype=np.float64), \
        (rng.standard_normal((300, 100)) * 0.5).astype(np.float32)
  if which in ('logistic25', 'logistic5', 'logistic30'):
    d = {'logistic25': 24, 'logistic5': 4, 'logistic30': 29}[which]
    n = {'logistic25': 1000, 'logistic5': 37, 'logistic30': 200}[which]
    X, y = tfp.targets.synthetic_logistic_data(n, d, seed=1)
    tg = tfp.targets.LogisticRegression(X, y)
    Xb = tg.features_with_bias
    return tg, otargets.LogisticRegression(Xb, y), \
        otargets.LogisticRegression(Xb.astype(np.float64), y.astype(np.float64), dtype=np.float64), \
        (0.3 * rng.standard_normal((130, d + 1))).astype(np.float32)
  if which in ('sv60', 'sv2516'):
    T = 60 if which == 'sv60' else 2516
    yv = tfp.targets.synthetic_sv_returns(T=T, seed=2)
    tg = tfp.targets.StochasticVolatility(yv)
    x = (0.3 * rng.standard_normal((9 if T > 100 else 33, T + 3))).astype(np.float32)
    x[:, 0] += 2.0
    x[:, 1] += 5.0
    return tg, otargets.StochasticVolatility(yv), \
        otargets.StochasticVolatility(yv.astype(np.float64), dtype=np.float64), x
  raise KeyError(which)


ALL_TARGETS = ['eight_schools', 'dense3', 'dense100', 'logistic5', 'logistic25', 'logistic30', 'sv60', 'sv2516']


@pytest.mark.parametrize('which', ALL_TARGETS)
def test_logp_grad_matches_oracle(tfp, which):
  tg, o32, o64, x = _targets(tfp, which)
  lp, g = tg.log_prob_and_grad(t(x))
  lp = lp.cpu().numpy(); g = g.cpu().numpy()
  lp64, g64 = o64.logp_grad(x.astype(np.float64))
  lp32, g32 = o32.logp_grad(x)
  # the kernel must be at least as close to the float64 truth as float32 arithmetic allows
  tol_lp = 5 * max(np.max(np.abs(lp32 - lp64) / (np.abs(lp64) + 1)), 2e-6)
  assert np.max(np.abs(lp - lp64) / (np.abs(lp64) + 1)) < tol_lp
  scale = np.max(np.abs(g64), axis=1, keepdims=True) + 1e-6
  tol_g = 5 * max(np.max(np.abs(g32 - g64) / scale), 2e-6)
  assert np.max(np.abs(g - g64) / scale) < tol_g


def test_target_is_the_target_log_prob_fn(tfp):
  tg, o32, _, x = _targets(tfp, 'eight_schools')
  lp = tg(t(x[:, 0]), t(x[:, 1]), t(x[:, 2:])).cpu().numpy()
  np.testing.assert_allclose(lp, o32.logp_grad(x)[0], rtol=2e-5, atol=1e-4)


def test_python_callable_target_is_rejected(tfp):
  with pytest.raises(TypeError):
    tfp.mcmc.HamiltonianMonteCarlo(lambda x: -x * x, step_size=0.1, num_leapfrog_steps=2)


# ------------------------------------------------------------------ leapfrog
@pytest.mark.parametrize('which,eps,L', [('eight_schools', 0.1, 3), ('dense3', 0.2, 5), ('dense100', 0.05, 3),
                                        ('logistic25', 0.02, 3), ('logistic5', 0.05, 4), ('sv60', 0.01, 3),
                                        ('sv2516', 0.005, 2)])
def test_leapfrog_trajectory_1e5(tfp, which, eps, L):
  """north_star: single leapfrog trajectories from identical (state, momentum) within 1e-5
  relative in float32 -- measured against the float64 oracle trajectory, whose float32
  counterpart carries the same order of rounding error."""
  tg, o32, o64, x = _targets(tfp, which)
  rng = np.random.default_rng(7)
  m = rng.standard_normal(x.shape).astype(np.float32)
  lp0, g0 = o32.logp_grad(x)
  from probability_b200 import _lib
  ctx = _lib.Context.get(dev()); ctx.bind_stream()
  B = x.shape[0]
  xm, xx, xl, xg = t(m), t(x), t(lp0), t(g0)
  step = torch.tensor([eps], device=dev())
  om, ox, og = torch.empty_like(xm), torch.empty_like(xx), torch.empty_like(xg)
  ol = torch.empty_like(xl)
  _lib.check(ctx.lib.pb2_leapfrog(ctx.handle, tg.handle(ctx), B, _lib.ptr(xm), _lib.ptr(xx), _lib.ptr(xl),
                                  _lib.ptr(xg), _lib.ptr(step), 0, L, _lib.ptr(om), _lib.ptr(ox), _lib.ptr(ol),
                                  _lib.ptr(og)), ctx.handle)
  lp64, g64 = o64.logp_grad(x.astype(np.float64))

  class T64:
    part_sizes = o64.part_sizes
    logp_grad = staticmethod(lambda z: o64.logp_grad(np.asarray(z, np.float64)))

  # float64 trajectory
  e = np.float64(eps)
  v = m.astype(np.float64) + 0.5 * e * g64
  xx64 = x.astype(np.float64)
  gg = g64
  for _ in range(L):
    xx64 = xx64 + e * v
    ll, gg = o64.logp_grad(xx64)
    v = v + e * gg
  m64 = v - 0.5 * e * gg
  r32 = omcmc.leapfrog(o32, m, x, lp0, g0, np.full(x.shape, eps, np.float32), L)
  base = max(rel_err(r32[1], xx64), 1e-6)
  print('leapfrog %s: rel err vs float64 trajectory: x %.2e (float32 oracle %.2e), m %.2e, lp %.2e' % (
      which, rel_err(ox.cpu().numpy(), xx64), base, rel_err(om.cpu().numpy(), m64), rel_err(ol.cpu().numpy(), ll)))
  assert rel_err(ox.cpu().numpy(), xx64) < max(1e-5, 3 * base)
  base_m = max(rel_err(r32[0], m64), 1e-6)
  assert rel_err(om.cpu().numpy(), m64) < max(1e-5, 3 * base_m)
  assert rel_err(ol.cpu().numpy(), ll) < max(1e-5, 3 * rel_err(r32[2], ll))


# ------------------------------------------------------------------ HMC
def _parts(tg, x):
  sizes = tg.part_sizes
  out, off = [], 0
  for n in sizes:
    piece = x[:, off:off + n]
    out.append(t(piece[:, 0] if n == 1 and len(sizes) > 1 else piece))
    off += n
  return out if len(out) > 1 else out[0]


def _flat(state):
  if isinstance(state, (list, tuple)):
    return torch.cat([s.reshape(s.shape[0], -1) for s in state], 1).cpu().numpy()
  return state.cpu().numpy()


@pytest.mark.parametrize('layout', [0, 1, 2])   # threefry partitionable / original, Philox4x32-10
@pytest.mark.parametrize('which,eps,L', [('eight_schools', 0.4, 3), ('dense3', 0.6, 4), ('logistic5', 0.2, 3),
                                        ('sv60', 0.02, 3)])
def test_hmc_one_step_matches_oracle(tfp, which, eps, L, layout):
  tfp.random.set_generator(('threefry', 'threefry_original', 'philox')[layout])
  try:
    tg, o32, _, x = _targets(tfp, which)
    state = _parts(tg, x)
    k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=eps, num_leapfrog_steps=L)
    pkr = k.bootstrap_results(state)
    seed = orng.key(11)
    new_state, kr = k.one_step(state, pkr, seed=seed)
    lp0, g0 = o32.logp_grad(x)
    ref = omcmc.hmc_one_step(o32, x, lp0, g0, eps, L, seed, layout)
    m0 = _flat(kr.proposed_results.initial_momentum)
    np.testing.assert_allclose(m0, ref['initial_momentum'], rtol=2e-6, atol=2e-7)   # same uint32 stream
    np.testing.assert_allclose(_flat(kr.proposed_state), ref['proposed_state'], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(kr.log_accept_ratio.cpu().numpy(), ref['log_accept_ratio'], rtol=1e-3, atol=2e-3)
    acc = kr.is_accepted.cpu().numpy()
    agree = acc == ref['is_accepted']
    assert agree.mean() >= 0.97
    got = _flat(new_state)
    np.testing.assert_allclose(got[agree], ref['state'][agree], rtol=1e-4, atol=1e-4)
    # accepted => state == proposed, rejected => unchanged (hmc_test.py:283-294)
    prop = _flat(kr.proposed_state)
    np.testing.assert_array_equal(got[acc], prop[acc])
    np.testing.assert_array_equal(got[~acc], x[~acc])
  finally:
    tfp.random.set_threefry_partitionable(True)


def test_hmc_fused_equals_composed_primitives(tfp):
  """HamiltonianMonteCarlo (one fused kernel) == MetropolisHastings(UncalibratedHMC) built from
  pb2_rng_normal + pb2_leapfrog + torch.where: same keys => same decisions."""
  tg, _, _, x = _targets(tfp, 'eight_schools')
  state = _parts(tg, x)
  seed = orng.key(23)
  k1 = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.3, num_leapfrog_steps=4)
  k2 = tfp.mcmc.MetropolisHastings(tfp.mcmc.UncalibratedHamiltonianMonteCarlo(tg, 0.3, 4))
  s1, r1 = k1.one_step(state, k1.bootstrap_results(state), seed=seed)
  s2, r2 = k2.one_step(state, k2.bootstrap_results(state), seed=seed)
  a1, a2 = r1.is_accepted.cpu().numpy(), r2.is_accepted.cpu().numpy()
  assert (a1 == a2).mean() > 0.99
  m = a1 == a2
  np.testing.assert_allclose(_flat(s1)[m], _flat(s2)[m], rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(r1.log_accept_ratio.cpu().numpy(), r2.log_accept_ratio.cpu().numpy(),
                             rtol=1e-4, atol=1e-4)


def test_hmc_nan_and_inf_reject(tfp):
  """Non-finite log-accept ratios reject (util.py:205-235; hmc_test.py:398-431,700-775):
  a huge step size drives Eight Schools to overflow."""
  tg, _, _, x = _targets(tfp, 'eight_schools')
  state = _parts(tg, x)
  k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=1e4, num_leapfrog_steps=3)
  s, r = k.one_step(state, k.bootstrap_results(state), seed=orng.key(1))
  lar = r.log_accept_ratio.cpu().numpy()
  acc = r.is_accepted.cpu().numpy()
  assert not np.isnan(lar).any()
  assert not acc[np.isneginf(lar)].any()
  np.testing.assert_array_equal(_flat(s)[~acc], x[~acc])
  assert np.isfinite(_flat(s)).all()


# ------------------------------------------------------------------ NUTS
@pytest.mark.parametrize('layout', [0, 1, 2])   # threefry partitionable / original, Philox4x32-10
@pytest.mark.parametrize('which,eps,depth', [('eight_schools', 0.3, 6), ('dense3', 0.5, 5), ('dense100', 0.3, 4),
                                            ('logistic5', 0.15, 5), ('logistic25', 0.03, 4), ('sv60', 0.03, 4)])
def test_nuts_one_step_matches_oracle(tfp, which, eps, depth, layout):
  tfp.random.set_generator(('threefry', 'threefry_original', 'philox')[layout])
  try:
    tg, o32, _, x = _targets(tfp, which)
    x = x[:64]
    state = _parts(tg, x)
    k = tfp.mcmc.NoUTurnSampler(tg, step_size=eps, max_tree_depth=depth)
    pkr = k.bootstrap_results(state)
    seed = orng.key(31)
    new_state, kr = k.one_step(state, pkr, seed=seed)
    lp0, g0 = o32.logp_grad(x)
    ref = omcmc.nuts_one_step(o32, x, lp0, g0, eps, seed, max_tree_depth=depth, layout=layout)
    nl = kr.leapfrogs_taken.cpu().numpy()
    same = nl == ref['leapfrogs_taken']
    # float rounding may flip a U-turn / multinomial comparison for a few chains (measured in round 2: 64 of 64
    # identical trees for every target and layout but one flip in one case)
    print('NUTS %s layout %d: identical trees %.4f' % (which, layout, same.mean()))
    assert same.mean() >= 0.97, (nl, ref['leapfrogs_taken'])
    got = _flat(new_state)
    close = np.isclose(got, ref['state'], rtol=2e-3, atol=2e-3).all(axis=1)
    print('   states close on identical trees %.4f' % close[same].mean())
    assert (close[same]).mean() >= 0.97
    for f in ('is_accepted', 'reach_max_depth', 'has_divergence'):
      assert (getattr(kr, f).cpu().numpy() == ref[f])[same & close].all(), f
    np.testing.assert_allclose(kr.log_accept_ratio.cpu().numpy()[same & close],
                               ref['log_accept_ratio'][same & close], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(kr.energy.cpu().numpy()[same & close], ref['energy'][same & close],
                               rtol=1e-3, atol=2e-3)
  finally:
    tfp.random.set_threefry_partitionable(True)


def test_nuts_divergence_and_max_depth_flags(tfp):
  tg, o32, _, x = _targets(tfp, 'eight_schools')
  state = _parts(tg, x[:64])
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=50.0, max_tree_depth=5)
  _, kr = k.one_step(state, k.bootstrap_results(state), seed=orng.key(2))
  assert kr.has_divergence.cpu().numpy().mean() > 0.5
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=1e-4, max_tree_depth=3)
  _, kr = k.one_step(state, k.bootstrap_results(state), seed=orng.key(2))
  assert kr.reach_max_depth.cpu().numpy().all()
  assert (kr.leapfrogs_taken.cpu().numpy() == 7).all()


# ------------------------------------------------------------------ sample_chain
def test_sample_chain_fused_equals_step_loop_and_thinning(tfp):
  """sample.py seed chaining: the fused driver == python loop over one_step; thinning picks every
  other state (hmc_test.py:91-140)."""
  tg, _, _, x = _targets(tfp, 'eight_schools')
  x = x[:32]
  state = _parts(tg, x)
  k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.3, num_leapfrog_steps=3)
  trace = lambda _, kr: (kr.is_accepted, kr.log_accept_ratio)
  s0, tr0 = tfp.mcmc.sample_chain(10, state, kernel=k, num_burnin_steps=3, trace_fn=trace, seed=5)
  s1, tr1 = tfp.mcmc.sample_chain(5, state, kernel=k, num_burnin_steps=3, num_steps_between_results=1,
                                  trace_fn=trace, seed=5)
  for a, b in zip(s0, s1):
    np.testing.assert_array_equal(a.cpu().numpy()[::2], b.cpu().numpy())
  np.testing.assert_array_equal(tr0[0].cpu().numpy()[::2], tr1[0].cpu().numpy())
  # python step loop (trace_fn computes on values -> not fusable)
  trace2 = lambda _, kr: (kr.is_accepted, kr.log_accept_ratio + 0.0)
  s2, tr2 = tfp.mcmc.sample_chain(10, state, kernel=k, num_burnin_steps=3, trace_fn=trace2, seed=5)
  for a, b in zip(s0, s2):
    np.testing.assert_array_equal(a.cpu().numpy(), b.cpu().numpy())
  np.testing.assert_array_equal(tr0[1].cpu().numpy(), tr2[1].cpu().numpy())


def test_sample_chain_matches_oracle_chain(tfp):
  tg, o32, _, x = _targets(tfp, 'eight_schools')
  x = x[:16]
  state = _parts(tg, x)
  k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.2, num_leapfrog_steps=2)
  s, tr = tfp.mcmc.sample_chain(4, state, kernel=k, num_burnin_steps=1, trace_fn=lambda _, kr: kr.is_accepted,
                                seed=17)
  ref_states, ref_trace, _ = omcmc.sample_chain(o32, 'hmc', x, 4, 1, 0, step_size=0.2, num_leapfrog_steps=2,
                                                seed=17)
  got = torch.cat([s[0][..., None], s[1][..., None], s[2]], -1).cpu().numpy()
  acc = tr.cpu().numpy()
  ref_acc = np.stack([r['is_accepted'] for r in ref_trace])
  ok = (acc == ref_acc).all(axis=0)     # chains whose decisions agree on every step
  assert ok.mean() > 0.8
  np.testing.assert_allclose(got[:, ok], ref_states[:, ok], rtol=1e-3, atol=1e-3)


def test_nuts_resume_from_final_kernel_results(tfp):
  """sample.py:60-78 CheckpointableStatesAndTrace: two halves == one run."""
  tg, _, _, x = _targets(tfp, 'dense3')
  state = t(x[:32])
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.5, max_tree_depth=5)
  tf_ = lambda _, kr: kr.leapfrogs_taken
  full = tfp.mcmc.sample_chain(6, state, kernel=k, trace_fn=tf_, seed=3)
  seed = tfp.random.sanitize_seed(3, salt='mcmc.sample_chain')
  # replay the seed chain by hand for the second half
  pkr = k.bootstrap_results(state)
  st = state
  for i in range(6):
    step_seed, seed = tfp.random.split_seed(seed)
    st, pkr = k.one_step(st, pkr, seed=step_seed)
    np.testing.assert_array_equal(st.cpu().numpy(), full.all_states[i].cpu().numpy())
    np.testing.assert_array_equal(pkr.leapfrogs_taken.cpu().numpy(), full.trace[i].cpu().numpy())


# ------------------------------------------------------------------ dual averaging pins
def test_dual_averaging_reference_pins(tfp):
  """dual_averaging_step_size_adaptation_test.py:51-57: step after two updates with a fake kernel
  whose log_accept_ratio is fixed."""
  import collections
  FakeMH = collections.namedtuple('FakeMH', 'accepted_results, log_accept_ratio')
  FakeAcc = collections.namedtuple('FakeAcc', 'step_size')

  class FakeKernel(tfp.mcmc.TransitionKernel):
    def __init__(self, step_size, log_accept_ratio, store_parameters_in_results=True):
      self._parameters = dict(step_size=step_size, log_accept_ratio=log_accept_ratio,
                              store_parameters_in_results=store_parameters_in_results)
    is_calibrated = True
    def one_step(self, state, pkr, seed=None):
      return state, FakeMH(pkr.accepted_results, t(np.asarray(self._parameters['log_accept_ratio'], np.float32)))
    def bootstrap_results(self, state):
      return FakeMH(FakeAcc(torch.tensor(self._parameters['step_size'], device=dev())),
                    t(np.asarray(self._parameters['log_accept_ratio'], np.float32)))

  pins = {0.7: 9.131008 / 10., 0.73: 9.642897 / 10., 0.74: 9.819825 / 10., 0.75: 1.0, 0.76: 10.183481 / 10.}
  for p, expect in pins.items():
    inner = FakeKernel(0.1, np.log(np.full(4, p)))
    k = tfp.mcmc.DualAveragingStepSizeAdaptation(inner, num_adaptation_steps=1)
    kr = k.bootstrap_results(torch.zeros(4, 1, device=dev()))
    for _ in range(2):
      _, kr = k.one_step(torch.zeros(4, 1, device=dev()), kr, seed=orng.key(0))
    # after step 1 (== num_adaptation_steps) the averaged step is used, then frozen
    got = float(kr.new_step_size)
    da = omcmc.DualAveraging(0.1, 1)
    da.update(np.log(np.full(4, p, np.float32))); da.update(np.log(np.full(4, p, np.float32)))
    np.testing.assert_allclose(got, float(da.step_size), rtol=1e-5)
  # exploration step after ONE update with many adaptation steps == the reference's _UPDATE_* pins
  for p, expect in pins.items():
    inner = FakeKernel(1.0, np.log(np.full(3, p)))
    k = tfp.mcmc.DualAveragingStepSizeAdaptation(inner, num_adaptation_steps=100)
    kr = k.bootstrap_results(torch.zeros(3, 1, device=dev()))
    _, kr = k.one_step(torch.zeros(3, 1, device=dev()), kr, seed=orng.key(0))
    np.testing.assert_allclose(float(kr.new_step_size), 10.0 * expect, rtol=2e-5)


def test_dual_averaging_fused_matches_oracle(tfp):
  tg, o32, _, x = _targets(tfp, 'eight_schools')
  x = x[:64]
  state = _parts(tg, x)
  k = tfp.mcmc.DualAveragingStepSizeAdaptation(
      tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.05, num_leapfrog_steps=3), num_adaptation_steps=8)
  res = tfp.mcmc.sample_chain(10, state, kernel=k, trace_fn=lambda _, kr: kr.inner_results.accepted_results.step_size,
                              seed=9, return_final_kernel_results=True)
  da = omcmc.DualAveraging(0.05, 8)
  _, ref_trace, _ = omcmc.sample_chain(o32, 'hmc', x, 10, 0, 0, step_size=0.05, num_leapfrog_steps=3, seed=9,
                                       dual_averaging=da)
  ref_steps = np.array([r['step_size'] for r in ref_trace], np.float32)
  got = res.trace.cpu().numpy()
  np.testing.assert_allclose(got[:4], ref_steps[:4], rtol=2e-3)   # later steps may see flipped decisions
  np.testing.assert_allclose(got, ref_steps, rtol=0.15)
  assert got[9] == got[8]                                          # frozen after num_adaptation_steps
  assert int(res.final_kernel_results.step) == 10
  # the step loop (one_step path) gives the same adaptation trace
  res2 = tfp.mcmc.sample_chain(10, state, kernel=k,
                               trace_fn=lambda _, kr: kr.inner_results.accepted_results.step_size + 0.0, seed=9)
  np.testing.assert_allclose(res2.trace.cpu().numpy(), got, rtol=1e-5)


# ------------------------------------------------------------------ diagnostics
def test_rhat_reference_pin(tfp):
  """diagnostic_test.py:405-420: arange(15).reshape(5,3) vs the NumPy formula."""
  state = np.arange(15.).reshape(5, 3).astype(np.float32)
  got = float(tfp.mcmc.potential_scale_reduction(t(state)).cpu())
  np.testing.assert_allclose(got, odiag.potential_scale_reduction(state), rtol=1e-5)
  n, m = 5., 3.
  b_div_n = np.var(state.mean(0), ddof=1)
  w = np.mean(np.var(state, axis=0, ddof=1))
  expect = ((m + 1) / m) * ((n - 1) / n * w + b_div_n) / w - (n - 1) / (m * n)
  np.testing.assert_allclose(got, expect, rtol=1e-5)


@pytest.mark.parametrize('split', [False, True])
def test_rhat_matches_oracle(tfp, split):
  rng = np.random.default_rng(0)
  x = rng.standard_normal((101, 6, 4)).astype(np.float32) + rng.standard_normal((1, 6, 4)).astype(np.float32)
  got = tfp.mcmc.potential_scale_reduction(t(x), split_chains=split).cpu().numpy()
  np.testing.assert_allclose(got, odiag.potential_scale_reduction(x, split_chains=split), rtol=2e-4)


@pytest.mark.parametrize('kw', [dict(), dict(filter_threshold=None, filter_beyond_positive_pairs=True),
                                dict(filter_beyond_lag=20, filter_threshold=None),
                                dict(cross_chain_dims=1), dict(cross_chain_dims=1, filter_beyond_positive_pairs=True,
                                                              filter_threshold=None)])
def test_ess_matches_oracle(tfp, kw):
  rng = np.random.default_rng(1)
  N, B, D = 400, 5, 3
  e = rng.standard_normal((N, B, D))
  x = np.zeros((N, B, D))
  for n in range(1, N):
    x[n] = 0.7 * x[n - 1] + e[n]
  x = x.astype(np.float32)
  got = tfp.mcmc.effective_sample_size(t(x), **kw).cpu().numpy()
  ref = odiag.effective_sample_size(x.astype(np.float64), **kw)
  np.testing.assert_allclose(got, ref, rtol=5e-3)


def test_ess_iid_is_n(tfp):
  """diagnostic_test.py:56-100: iid N(0,1), ESS ~= N."""
  rng = np.random.default_rng(2)
  x = rng.standard_normal((5000, 4)).astype(np.float32)
  got = tfp.mcmc.effective_sample_size(t(x), filter_threshold=0.).cpu().numpy()
  np.testing.assert_allclose(got, 5000., rtol=0.1)
  got = tfp.mcmc.effective_sample_size(t(x), filter_threshold=None, filter_beyond_positive_pairs=True).cpu().numpy()
  np.testing.assert_allclose(got, 5000., rtol=0.25)


# ------------------------------------------------------------------ tcgen05 tile kernels (dense Gaussian, B >= 256)
def _dense100_state(B, seed=0):
  import probability_b200 as tfp_
  tg = tfp_.targets.IllConditionedGaussian()
  rng = np.random.default_rng(seed)
  L = np.linalg.cholesky(tg.covariance)
  x0 = (rng.standard_normal((B, 100)) @ L.T).astype(np.float32)
  return tg, otargets.DenseGaussian(tg.precision, tg.log_normalizer), x0


def test_tensor_core_logp_grad_is_fp32_accurate(tfp):
  """pb2_dense_logp_grad_tc (tcgen05, 3xTF32): as close to float64 as the FP32-FMA kernel."""
  from probability_b200 import _lib
  tg, o32, x0 = _dense100_state(300)
  ctx = _lib.Context.get(dev()); ctx.bind_stream()
  xt = t(x0)
  lp = torch.empty(300, device=dev()); g = torch.empty(300, 100, device=dev())
  _lib.check(ctx.lib.pb2_dense_logp_grad_tc(ctx.handle, tg.handle(ctx), 300, _lib.ptr(xt), _lib.ptr(lp), _lib.ptr(g)),
             ctx.handle)
  g64 = -(x0.astype(np.float64) @ tg.precision.astype(np.float64))
  lp64 = 0.5 * np.sum(x0 * g64, 1) + tg.log_normalizer
  lp32, g32 = o32.logp_grad(x0)
  scale = np.abs(g64).max(1, keepdims=True)
  assert np.max(np.abs(g.cpu().numpy() - g64) / scale) < 3 * max(np.max(np.abs(g32 - g64) / scale), 1e-6)
  assert np.max(np.abs(lp.cpu().numpy() - lp64) / np.abs(lp64)) < 3 * max(np.max(np.abs(lp32 - lp64) / np.abs(lp64)), 1e-6)


@pytest.mark.parametrize('variant', [0, 1])
def test_tile_hmc_matches_oracle(tfp, variant):
  from probability_b200 import _lib
  tg, o32, x0 = _dense100_state(300)
  ctx = _lib.Context.get(dev())
  ctx.set_int('dense_variant', variant)
  try:
    k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.5, num_leapfrog_steps=5)
    seed = orng.key(4)
    s, r = k.one_step(t(x0), k.bootstrap_results(t(x0)), seed=seed)
    lp0, g0 = o32.logp_grad(x0)
    ref = omcmc.hmc_one_step(o32, x0, lp0, g0, 0.5, 5, seed)
    acc = r.is_accepted.cpu().numpy()
    agree = acc == ref['is_accepted']
    assert agree.mean() > 0.98
    np.testing.assert_allclose(r.proposed_results.initial_momentum.cpu().numpy(), ref['initial_momentum'], atol=1e-6)
    np.testing.assert_allclose(s.cpu().numpy()[agree], ref['state'][agree], rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(r.log_accept_ratio.cpu().numpy(), ref['log_accept_ratio'], atol=2e-2)
  finally:
    ctx.set_int('dense_variant', 0)


@pytest.mark.parametrize('depth,eps', [(3, 0.3), (6, 0.5)])
def test_tile_nuts_matches_oracle_and_warp_kernel(tfp, depth, eps):
  """The lock-step 128-chain tile kernel IS the reference's batched algorithm: same trees as the
  oracle, and as the warp-per-chain kernel (per-chain early exit)."""
  from probability_b200 import _lib
  tg, o32, x0 = _dense100_state(300)
  ctx = _lib.Context.get(dev())
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=eps, max_tree_depth=depth)
  seed = orng.key(4)
  outs = {}
  try:
    for variant in (0, 1):
      ctx.set_int('dense_variant', variant)
      s, r = k.one_step(t(x0), k.bootstrap_results(t(x0)), seed=seed)
      outs[variant] = (s.cpu().numpy(), r)
  finally:
    ctx.set_int('dense_variant', 0)
  lp0, g0 = o32.logp_grad(x0)
  ref = omcmc.nuts_one_step(o32, x0, lp0, g0, eps, seed, max_tree_depth=depth)
  for variant in (0, 1):
    s, r = outs[variant]
    same = r.leapfrogs_taken.cpu().numpy() == ref['leapfrogs_taken']
    assert same.mean() >= 0.97
    close = np.isclose(s, ref['state'], rtol=2e-3, atol=2e-3).all(1)
    assert close[same].mean() >= 0.97
    for f in ('is_accepted', 'reach_max_depth', 'has_divergence'):
      assert (getattr(r, f).cpu().numpy() == ref[f])[same & close].all(), f
  assert (outs[0][1].leapfrogs_taken == outs[1][1].leapfrogs_taken).float().mean() > 0.97


@pytest.mark.parametrize('B,depth', [(1000, 7), (4096, 10), (10000, 10)])
def test_tile_nuts_async_lanes_are_bit_identical_to_lockstep(tfp, B, depth):
  """NUTS on the dense target runs the asynchronous-lane tile kernel (pb2_tile_nuts.cu): every lane of a
  64-chain tile is at its own leaf of its own tree / transition, chains go back to a FIFO after every
  transition.  Because every random draw is keyed by (step seed, global chain index) and each chain's
  arithmetic is independent of its tile mates, the result must equal the lock-step tile kernel's (the
  reference's literal batched algorithm) bit for bit -- ragged B (partially filled tiles) and more chains
  than lanes (10000 > 148 * 64: chains queue for lanes) included -- and must not depend on scheduling."""
  from probability_b200 import _lib
  tg, _, x0 = _dense100_state(B, seed=3)
  ctx = _lib.Context.get(dev())
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=depth)
  fields = lambda _, kr: (kr.leapfrogs_taken, kr.is_accepted, kr.energy, kr.target_log_prob,
                          kr.has_divergence, kr.reach_max_depth, kr.log_accept_ratio)
  outs = {}
  try:
    for name, variant in (('async', 0), ('lockstep', 3), ('async_again', 0)):
      ctx.set_int('dense_variant', variant)
      res = tfp.mcmc.sample_chain(5, t(x0), kernel=k, trace_fn=fields, seed=7)
      outs[name] = [res.all_states.cpu().numpy()] + [f.cpu().numpy() for f in res.trace]
  finally:
    ctx.set_int('dense_variant', 0)
  for a, b in zip(outs['async'], outs['lockstep']):
    np.testing.assert_array_equal(a, b)
  for a, b in zip(outs['async'], outs['async_again']):
    np.testing.assert_array_equal(a, b)
  assert np.isfinite(outs['async'][0]).all() and np.isfinite(outs['async'][3]).all()
  lf = outs['async'][1]
  assert lf.min() >= 1 and lf.max() <= 2 ** depth - 1 and len(np.unique(lf)) > 3


def test_tile_nuts_async_single_transitions_equal_fused_run(tfp):
  """A python loop over one_step (one launch per transition: the path of a per-step host loop and of step-size
  adaptation) == one fused K-transition launch, on the asynchronous-lane kernel."""
  tg, _, x0 = _dense100_state(700, seed=4)
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.7, max_tree_depth=8)
  fused = tfp.mcmc.sample_chain(4, t(x0), kernel=k, trace_fn=lambda _, kr: kr.leapfrogs_taken, seed=11)
  # a trace_fn that computes on values cannot be fused: sample_chain falls back to its one_step loop
  loop = tfp.mcmc.sample_chain(4, t(x0), kernel=k, trace_fn=lambda _, kr: kr.leapfrogs_taken + 0, seed=11)
  np.testing.assert_array_equal(loop.all_states.cpu().numpy(), fused.all_states.cpu().numpy())
  np.testing.assert_array_equal(loop.trace.cpu().numpy(), fused.trace.cpu().numpy())


@pytest.mark.parametrize('case', ['unrolled2', 'per_chain_step'])
def test_tile_nuts_options_match_lockstep_and_oracle(tfp, case):
  """unrolled_leapfrog_steps > 1 (nuts.py:826-843: L leapfrogs per tree leaf) and per-chain step sizes on the
  tensor-core tile kernels: asynchronous lanes == lock-step bit for bit, and the trees match the oracle's."""
  from probability_b200 import _lib
  B = 640
  tg, o32, x0 = _dense100_state(B, seed=6)
  ctx = _lib.Context.get(dev())
  if case == 'unrolled2':
    eps_np, kw = np.float32(0.35), dict(unrolled_leapfrog_steps=2)
    k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.35, max_tree_depth=7, **kw)
  else:
    eps_np = (0.4 + 0.4 * np.random.default_rng(1).random((B, 1))).astype(np.float32)
    kw = {}
    k = tfp.mcmc.NoUTurnSampler(tg, step_size=t(eps_np), max_tree_depth=7)
  seed = orng.key(21)
  outs = {}
  try:
    for variant in (0, 3):
      ctx.set_int('dense_variant', variant)
      s, r = k.one_step(t(x0), k.bootstrap_results(t(x0)), seed=seed)
      outs[variant] = (s.cpu().numpy(), r.leapfrogs_taken.cpu().numpy(), r.energy.cpu().numpy())
  finally:
    ctx.set_int('dense_variant', 0)
  for a, b in zip(outs[0], outs[3]):
    np.testing.assert_array_equal(a, b)
  lp0, g0 = o32.logp_grad(x0)
  ref = omcmc.nuts_one_step(o32, x0, lp0, g0, eps_np, seed, max_tree_depth=7, **kw)
  same = outs[0][1] == ref['leapfrogs_taken']
  assert same.mean() >= 0.97
  close = np.isclose(outs[0][0], ref['state'], rtol=2e-3, atol=2e-3).all(1)
  assert close[same].mean() >= 0.97


@pytest.mark.parametrize('sampler', ['nuts', 'hmc'])
def test_tile_kernels_chain_sharding_is_bit_identical(tfp, sampler):
  """Chains sharded over ranks: the RNG counters use the GLOBAL chain index and the global batch size,
  so two half-batches run with ChainShard(offset, total) reproduce the unsharded run bit for bit -- here on the
  tensor-core tile kernels, both shards on one GPU."""
  B = 1024
  tg, _, x0 = _dense100_state(B, seed=8)

  def run(x, shard):
    if sampler == 'nuts':
      k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.7, max_tree_depth=8, experimental_chain_shard=shard)
      tr = lambda _, kr: (kr.leapfrogs_taken, kr.energy)
    else:
      k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.5, num_leapfrog_steps=6, experimental_chain_shard=shard)
      tr = lambda _, kr: (kr.is_accepted, kr.log_accept_ratio)
    r = tfp.mcmc.sample_chain(3, t(x), kernel=k, trace_fn=tr, seed=13)
    return [r.all_states.cpu().numpy()] + [f.cpu().numpy() for f in r.trace]

  full = run(x0, None)
  lo = run(x0[:B // 2], tfp.mcmc.ChainShard(0, B))
  hi = run(x0[B // 2:], tfp.mcmc.ChainShard(B // 2, B))
  for f, a, b in zip(full, lo, hi):
    np.testing.assert_array_equal(f, np.concatenate([a, b], axis=1))


@pytest.mark.parametrize('B,N', [(300, 1000), (128, 77)])
def test_tensor_core_logistic_logp_grad_is_fp32_accurate(tfp, B, N):
  """pb2_logistic_logp_grad_tc (two tcgen05 3xTF32 contractions around the sigmoid, logits kept in TMEM):
  as close to float64 as the FP32 warp-per-chain kernel."""
  from probability_b200 import _lib
  X, y = otargets.synthetic_logistic_data(N, 24, seed=3)        # X includes the ones column
  tg = tfp.targets.LogisticRegression(X[:, :-1], y)
  o64 = otargets.LogisticRegression(X, y, dtype=np.float64)
  th = (0.5 * np.random.default_rng(5).standard_normal((B, 25))).astype(np.float32)
  ctx = _lib.Context.get(dev()); ctx.bind_stream()
  tht = t(th)
  lp = torch.empty(B, device=dev()); g = torch.empty(B, 25, device=dev())
  _lib.check(ctx.lib.pb2_logistic_logp_grad_tc(ctx.handle, tg.handle(ctx), B, _lib.ptr(tht), _lib.ptr(lp), _lib.ptr(g)),
             ctx.handle)
  lp2 = torch.empty(B, device=dev()); g2 = torch.empty(B, 25, device=dev())
  _lib.check(ctx.lib.pb2_logp_grad(ctx.handle, tg.handle(ctx), B, _lib.ptr(tht), _lib.ptr(lp2), _lib.ptr(g2)), ctx.handle)
  lp64, g64 = o64.logp_grad(th.astype(np.float64))
  scale = np.abs(g64).max(1, keepdims=True)
  err_tc = np.max(np.abs(g.cpu().numpy() - g64) / scale)
  err_fp = np.max(np.abs(g2.cpu().numpy() - g64) / scale)
  assert err_tc < 3 * max(err_fp, 2e-6), (err_tc, err_fp)
  elp_tc = np.max(np.abs(lp.cpu().numpy() - lp64) / np.abs(lp64))
  elp_fp = np.max(np.abs(lp2.cpu().numpy() - lp64) / np.abs(lp64))
  assert elp_tc < 3 * max(elp_fp, 2e-6), (elp_tc, elp_fp)


@pytest.mark.parametrize('D', [50, 77, 100])
def test_tile_kernels_general_dense_gaussian_with_loc(tfp, D):
  """The tensor-core tile kernels on a general dense Gaussian -- non-zero location, dimension not a multiple of the
  13-dim thread segments (the padded columns must stay out of the contraction): gradient vs float64, HMC and NUTS vs
  the oracle, asynchronous lanes == lock-step."""
  from probability_b200 import _lib
  rng = np.random.default_rng(D)
  A = rng.standard_normal((D, D))
  cov = A @ A.T / D + 0.3 * np.eye(D)
  loc = (3.0 * rng.standard_normal(D)).astype(np.float32)
  tg = tfp.targets.DenseGaussian(covariance=cov, loc=loc)
  o32 = otargets.DenseGaussian(tg.precision, tg.log_normalizer, loc=loc)
  B = 320
  x0 = (loc + rng.standard_normal((B, D)) @ np.linalg.cholesky(cov).T).astype(np.float32)
  ctx = _lib.Context.get(dev()); ctx.bind_stream()
  # gradient primitive (tcgen05) vs float64
  lp = torch.empty(B, device=dev()); g = torch.empty(B, D, device=dev())
  _lib.check(ctx.lib.pb2_dense_logp_grad_tc(ctx.handle, tg.handle(ctx), B, _lib.ptr(t(x0)), _lib.ptr(lp), _lib.ptr(g)),
             ctx.handle)
  g64 = -((x0.astype(np.float64) - loc) @ tg.precision.astype(np.float64))
  assert np.max(np.abs(g.cpu().numpy() - g64)) / np.abs(g64).max() < 2e-6
  seed = orng.key(5)
  lp0, g0 = o32.logp_grad(x0)
  # HMC tile kernel vs oracle
  k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.25, num_leapfrog_steps=4)
  s, r = k.one_step(t(x0), k.bootstrap_results(t(x0)), seed=seed)
  ref = omcmc.hmc_one_step(o32, x0, lp0, g0, 0.25, 4, seed)
  agree = r.is_accepted.cpu().numpy() == ref['is_accepted']
  assert agree.mean() > 0.97
  np.testing.assert_allclose(s.cpu().numpy()[agree], ref['state'][agree], rtol=1e-4, atol=2e-3)
  # NUTS: asynchronous lanes == lock-step bit for bit, and the oracle's trees
  kn = tfp.mcmc.NoUTurnSampler(tg, step_size=0.3, max_tree_depth=7)
  outs = {}
  try:
    for variant in (0, 3):
      ctx.set_int('dense_variant', variant)
      sn, rn = kn.one_step(t(x0), kn.bootstrap_results(t(x0)), seed=seed)
      outs[variant] = (sn.cpu().numpy(), rn.leapfrogs_taken.cpu().numpy())
  finally:
    ctx.set_int('dense_variant', 0)
  np.testing.assert_array_equal(outs[0][0], outs[3][0])
  np.testing.assert_array_equal(outs[0][1], outs[3][1])
  refn = omcmc.nuts_one_step(o32, x0, lp0, g0, 0.3, seed, max_tree_depth=7)
  same = outs[0][1] == refn['leapfrogs_taken']
  assert same.mean() >= 0.97
  close = np.isclose(outs[0][0], refn['state'], rtol=2e-3, atol=2e-3).all(1)
  assert close[same].mean() >= 0.97


def test_tile_nuts_async_burnin_and_thinning(tfp):
  """sample.py:359-366 emission schedule on the asynchronous-lane kernel: burn-in steps are not emitted, thinning
  picks every other state (hmc_test.py:91-140), traced fields line up with the states."""
  tg, _, x0 = _dense100_state(384, seed=9)
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.7, max_tree_depth=7)
  tr = lambda _, kr: (kr.leapfrogs_taken, kr.target_log_prob)
  a = tfp.mcmc.sample_chain(8, t(x0), kernel=k, num_burnin_steps=3, trace_fn=tr, seed=31)
  b = tfp.mcmc.sample_chain(4, t(x0), kernel=k, num_burnin_steps=3, num_steps_between_results=1, trace_fn=tr, seed=31)
  np.testing.assert_array_equal(a.all_states.cpu().numpy()[::2], b.all_states.cpu().numpy())
  np.testing.assert_array_equal(a.trace[0].cpu().numpy()[::2], b.trace[0].cpu().numpy())
  np.testing.assert_array_equal(a.trace[1].cpu().numpy()[::2], b.trace[1].cpu().numpy())
  # the traced log-prob is the log-prob of the emitted state
  o32 = otargets.DenseGaussian(tg.precision, tg.log_normalizer)
  lp, _ = o32.logp_grad(a.all_states[-1].cpu().numpy())
  np.testing.assert_allclose(lp, a.trace[1][-1].cpu().numpy(), rtol=2e-4, atol=5e-2)


def test_logistic_hmc_on_tensor_cores_matches_oracle_and_fp32_kernel(tfp):
  """T3's tensor-core transition path: HamiltonianMonteCarlo on LogisticRegression(tensor_core_transitions=True) runs the
  leapfrogs in lock-step with the tcgen05 log-prob + gradient (pb2_logistic_tc_leapfrog).  Same momentum stream, same
  decisions as the oracle and as the FP32 warp-per-chain kernel (float rounding may flip a comparison)."""
  X, y = tfp.targets.synthetic_logistic_data(1000, 24, seed=1)
  tc = tfp.targets.LogisticRegression(X, y, tensor_core_transitions=True)
  fp = tfp.targets.LogisticRegression(X, y)
  o32 = otargets.LogisticRegression(fp.features_with_bias, y)
  x0 = (0.3 * np.random.default_rng(5).standard_normal((300, 25))).astype(np.float32)
  seed = orng.key(17)
  lp0, g0 = o32.logp_grad(x0)
  ref = omcmc.hmc_one_step(o32, x0, lp0, g0, 0.03, 6, seed)
  outs = {}
  for name, tg in (('tc', tc), ('fp32', fp)):
    k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.03, num_leapfrog_steps=6)
    s, r = k.one_step(t(x0), k.bootstrap_results(t(x0)), seed=seed)
    outs[name] = (s.cpu().numpy(), r)
    acc = r.is_accepted.cpu().numpy()
    agree = acc == ref['is_accepted']
    assert agree.mean() >= 0.97, name
    np.testing.assert_allclose(r.proposed_state.cpu().numpy(), ref['proposed_state'], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(r.log_accept_ratio.cpu().numpy(), ref['log_accept_ratio'], rtol=2e-3, atol=3e-3)
    np.testing.assert_allclose(s.cpu().numpy()[agree], ref['state'][agree], rtol=2e-4, atol=2e-4)
  np.testing.assert_allclose(outs['tc'][1].proposed_results.initial_momentum.cpu().numpy(),
                             outs['fp32'][1].proposed_results.initial_momentum.cpu().numpy(), rtol=0, atol=0)
  # a few transitions through sample_chain (the step loop drives the lock-step leapfrog) stay in agreement
  k = tfp.mcmc.HamiltonianMonteCarlo(tc, step_size=0.03, num_leapfrog_steps=6)
  kf = tfp.mcmc.HamiltonianMonteCarlo(fp, step_size=0.03, num_leapfrog_steps=6)
  a = tfp.mcmc.sample_chain(4, t(x0), kernel=k, seed=3, trace_fn=lambda _, kr: kr.is_accepted)
  b = tfp.mcmc.sample_chain(4, t(x0), kernel=kf, seed=3, trace_fn=lambda _, kr: kr.is_accepted)
  assert (a.trace == b.trace).float().mean() > 0.97
