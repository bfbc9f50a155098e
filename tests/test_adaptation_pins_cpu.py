"""CPU tests of the step-size adaptation host logic against the reference's own pins, restated with the reference's
fake kernels (tfp/mcmc/dual_averaging_step_size_adaptation_test.py:60-163): the adaptation classes only see the
kernel results, so the general (non-fused) update runs on CPU tensors and no GPU kernel is involved.

  * `_UPDATE_*` and the 8 broadcasting cases              dual_averaging..._test.py:51-57, 605-671
  * testListStep / testWrapped / testChainLogProb*        :177-290
  * NaN accept prob, target_accept_prob validation        :224-240, 292-311
  * SimpleStepSizeAdaptation: x(1+rate) / /(1+rate), list steps, finite adaptation
                                                          simple_step_size_adaptation_test.py, hmc_test.py:917-952
"""
import collections

import numpy as np
import pytest
import torch

import probability_b200 as tfp
from probability_b200.mcmc import dual_averaging_step_size_adaptation as duassa
from probability_b200.mcmc import simple_step_size_adaptation as sssa
from probability_b200.mcmc.kernel import TransitionKernel

_INITIAL_T = 10.0
_EXPLORATION_SHRINKAGE = 0.05
_UPDATE_M05 = 9.131008  # err = -0.05
_UPDATE_M02 = 9.642897  # err = -0.02
_UPDATE_M01 = 9.819825  # err = -0.01
_UPDATE_0 = 10.  # err = 0
_UPDATE_01 = 10.183481  # err = +0.01

FakeMHKernelResults = collections.namedtuple('FakeMHKernelResults', 'accepted_results, log_accept_ratio')
FakeSteppedKernelResults = collections.namedtuple('FakeSteppedKernelResults', 'step_size')
FakeWrapperKernelResults = collections.namedtuple('FakeWrapperKernelResults', 'inner_results')


def _t(v):
  if isinstance(v, (list, tuple)):
    return [_t(u) for u in v]
  return torch.as_tensor(np.asarray(v, np.float32))


class FakeSteppedKernel(TransitionKernel):
  def __init__(self, step_size, store_parameters_in_results=False, experimental_shard_axis_names=None):
    self._parameters = dict(step_size=step_size, store_parameters_in_results=store_parameters_in_results,
                            experimental_shard_axis_names=experimental_shard_axis_names)

  def one_step(self, current_state, previous_kernel_results, seed=None):
    return current_state, previous_kernel_results

  def bootstrap_results(self, current_state):
    return FakeSteppedKernelResults(step_size=_t(self._parameters['step_size']))

  @property
  def experimental_shard_axis_names(self):
    return self._parameters['experimental_shard_axis_names']

  def experimental_with_shard_axes(self, shard_axes):
    return self.copy(experimental_shard_axis_names=shard_axes)

  is_calibrated = False


class FakeMHKernel(TransitionKernel):
  def __init__(self, inner_kernel, log_accept_ratio, store_parameters_in_results=False):
    self._parameters = dict(inner_kernel=inner_kernel, log_accept_ratio=log_accept_ratio,
                            store_parameters_in_results=store_parameters_in_results)

  inner_kernel = property(lambda self: self._parameters['inner_kernel'])

  def one_step(self, current_state, previous_kernel_results, seed=None):
    new_state, new_acc = self.inner_kernel.one_step(current_state, previous_kernel_results.accepted_results,
                                                    seed=seed)
    return new_state, previous_kernel_results._replace(accepted_results=new_acc)

  def bootstrap_results(self, current_state):
    return FakeMHKernelResults(accepted_results=self.inner_kernel.bootstrap_results(current_state),
                               log_accept_ratio=_t(self._parameters['log_accept_ratio']))

  @property
  def experimental_shard_axis_names(self):
    return self.inner_kernel.experimental_shard_axis_names

  def experimental_with_shard_axes(self, shard_axes):
    return self.copy(inner_kernel=self.inner_kernel.experimental_with_shard_axes(shard_axes))

  is_calibrated = True


class FakeWrapperKernel(TransitionKernel):
  def __init__(self, inner_kernel):
    self._parameters = dict(inner_kernel=inner_kernel)

  inner_kernel = property(lambda self: self._parameters['inner_kernel'])

  def one_step(self, current_state, previous_kernel_results, seed=None):
    new_state, new_inner = self.inner_kernel.one_step(current_state, previous_kernel_results.inner_results)
    return new_state, previous_kernel_results._replace(inner_results=new_inner)

  def bootstrap_results(self, current_state):
    return FakeWrapperKernelResults(inner_results=self.inner_kernel.bootstrap_results(current_state))

  @property
  def is_calibrated(self):
    return self.inner_kernel.is_calibrated


def _two_steps(kernel, state):
  kr = kernel.bootstrap_results(state)
  for _ in range(2):
    _, kr = kernel.one_step(state, kr)
  return kr


# ------------------------------------------------------------------------------------------- dual averaging
def test_store_parameters_turned_on():
  kernel = FakeWrapperKernel(FakeSteppedKernel(step_size=0.5))
  assert not kernel.inner_kernel.parameters['store_parameters_in_results']
  kernel = tfp.mcmc.DualAveragingStepSizeAdaptation(kernel, num_adaptation_steps=1, validate_args=True)
  assert kernel.inner_kernel.inner_kernel.parameters['store_parameters_in_results']


def test_list_step():
  kernel = FakeMHKernel(FakeSteppedKernel(step_size=np.float32([0.1, 0.2, 0.3])),
                        log_accept_ratio=np.log([0.74, 0.76, 0.76]))
  kernel = tfp.mcmc.DualAveragingStepSizeAdaptation(kernel, num_adaptation_steps=1, validate_args=True)
  kr = _two_steps(kernel, torch.zeros(3))
  expected = np.exp(np.log(10. * np.array([0.1, 0.2, 0.3])) -
                    np.array([0.01, -0.01, -0.01]) / ((_INITIAL_T + 1.) * _EXPLORATION_SHRINKAGE))
  np.testing.assert_allclose(kr.inner_results.accepted_results.step_size.numpy(), expected, rtol=1e-5)


def test_wrapped():
  kernel = FakeWrapperKernel(FakeMHKernel(FakeSteppedKernel(step_size=0.1), log_accept_ratio=np.log(0.76)))
  kernel = tfp.mcmc.DualAveragingStepSizeAdaptation(kernel, num_adaptation_steps=1, validate_args=True)
  kr = _two_steps(kernel, torch.tensor(0.))
  expected = np.exp(np.log(10. * 0.1) - -0.01 / ((_INITIAL_T + 1.) * _EXPLORATION_SHRINKAGE))
  np.testing.assert_allclose(kr.inner_results.inner_results.accepted_results.step_size.numpy(), expected, rtol=1e-5)


def test_recovers_from_nan_accept_prob():
  kernel = FakeMHKernel(FakeSteppedKernel(step_size=0.1), log_accept_ratio=np.nan)
  kernel = tfp.mcmc.DualAveragingStepSizeAdaptation(kernel, num_adaptation_steps=1, validate_args=True)
  kr = _two_steps(kernel, torch.tensor(0.))
  assert np.isfinite(kr.inner_results.accepted_results.step_size.numpy())


@pytest.mark.parametrize('target,errs', [(None, [0.01, -0.01]), ([0.7, 0.8], [-0.04, 0.04])])
def test_chain_log_prob_targets(target, errs):
  init_step = np.float32([0.1, 0.2])
  kernel = FakeMHKernel(FakeSteppedKernel(step_size=init_step), log_accept_ratio=np.log([0.74, 0.76]))
  kw = {} if target is None else dict(target_accept_prob=_t(target))
  kernel = tfp.mcmc.DualAveragingStepSizeAdaptation(
      kernel, num_adaptation_steps=1, validate_args=True,
      log_accept_prob_getter_fn=lambda pkr: torch.minimum(torch.zeros(()), pkr.log_accept_ratio), **kw)
  kr = _two_steps(kernel, torch.zeros(2))
  expected = np.exp(np.log(10. * init_step) - np.array(errs) / ((_INITIAL_T + 1.) * _EXPLORATION_SHRINKAGE))
  np.testing.assert_allclose(kr.inner_results.accepted_results.step_size.numpy(), expected, rtol=1e-5)


@pytest.mark.parametrize('tap,message', [(-1., '`target_accept_prob` must be > 0.'),
                                         (0., '`target_accept_prob` must be > 0.'), (0.999, None),
                                         (1., '`target_accept_prob` must be < 1.')])
def test_target_accept_prob_checks(tap, message):
  def impl():
    kernel = FakeMHKernel(FakeSteppedKernel(step_size=1.), log_accept_ratio=0.)
    kernel = tfp.mcmc.DualAveragingStepSizeAdaptation(kernel, num_adaptation_steps=1, target_accept_prob=tap,
                                                      validate_args=True)
    kernel.bootstrap_results(torch.zeros(()))
  if message:
    with pytest.raises(ValueError, match=message.replace('`', '.')):
      impl()
  else:
    impl()


_BROADCAST_CASES = [
    (1., _UPDATE_M01),
    ([1., np.ones([3, 1])], [_UPDATE_M01, np.array([[_UPDATE_M02], [_UPDATE_01], [_UPDATE_M02]])]),
    ([1., np.ones([2, 3, 1])], [_UPDATE_M01, np.array([[[_UPDATE_M05], [_UPDATE_01], [_UPDATE_M02]],
                                                       [[_UPDATE_01], [_UPDATE_01], [_UPDATE_M02]]])]),
    ([1., np.ones([2, 1, 1])], [_UPDATE_M01, np.array([[[_UPDATE_M02]], [[_UPDATE_0]]])]),
    ([1., np.ones([1, 3, 1])], [_UPDATE_M01, np.array([[[_UPDATE_M02], [_UPDATE_01], [_UPDATE_M02]]])]),
    ([1., np.ones([1, 1, 1])], [_UPDATE_M01, np.array([[[_UPDATE_M01]]])]),
    ([1., np.ones([1, 1])], [_UPDATE_M01, np.array([[_UPDATE_M01]])]),
    ([1., np.ones([1])], [_UPDATE_M01, np.array([_UPDATE_M01])]),
]


@pytest.mark.parametrize('old_step_size,new_step_size', _BROADCAST_CASES)
def test_broadcasting(old_step_size, new_step_size):
  """The 8 cases of DualAveragingStepSizeAdaptationStaticBroadcastingTest.testBroadcasting (:605-671)."""
  log_accept_ratio = np.log([[0.70, 0.76, 0.73], [0.76, 0.76, 0.73]])
  state = [torch.zeros(2, 3), torch.zeros(2, 3, 4)]
  kernel = FakeMHKernel(FakeSteppedKernel(step_size=old_step_size), log_accept_ratio=log_accept_ratio)
  kernel = tfp.mcmc.DualAveragingStepSizeAdaptation(kernel, target_accept_prob=0.75, num_adaptation_steps=1,
                                                    validate_args=True)
  kr = _two_steps(kernel, state)
  got = kr.inner_results.accepted_results.step_size
  if isinstance(new_step_size, list):
    assert isinstance(got, list) and len(got) == len(new_step_size)
    for g, e in zip(got, new_step_size):
      assert tuple(g.shape) == np.shape(e)
      np.testing.assert_allclose(g.numpy(), e, rtol=2e-6)
  else:
    np.testing.assert_allclose(got.numpy(), new_step_size, rtol=2e-6)


def test_error_sum_keeps_accumulating_after_adaptation():
  # :437-439: error_sum is updated on every step; step size and averaging step are frozen after num_adaptation_steps
  kernel = FakeMHKernel(FakeSteppedKernel(step_size=1.), log_accept_ratio=np.log([0.5, 0.5]))
  kernel = tfp.mcmc.DualAveragingStepSizeAdaptation(kernel, num_adaptation_steps=2)
  state = torch.zeros(2)
  kr = kernel.bootstrap_results(state)
  steps = []
  for _ in range(5):
    _, kr = kernel.one_step(state, kr)
    steps.append(float(kr.new_step_size))
  np.testing.assert_allclose(float(kr.error_sum[0]), 5 * 0.25, rtol=1e-6)
  assert steps[1] == steps[2] == steps[4]
  assert int(kr.step) == 5


def test_shard_axes_forwarded():
  kernel = FakeMHKernel(FakeSteppedKernel(step_size=1.), log_accept_ratio=0.)
  kernel = tfp.mcmc.DualAveragingStepSizeAdaptation(kernel, num_adaptation_steps=1)
  sharded = kernel.experimental_with_shard_axes(['foo'])
  assert sharded.inner_kernel.inner_kernel.experimental_shard_axis_names == ['foo']
  assert sharded.experimental_shard_axis_names == ['foo']


# ------------------------------------------------------------------------------------ simple step-size adaptation
@pytest.mark.parametrize('p,up', [(0.76, True), (0.74, False)])
def test_simple_adaptation_direction(p, up):
  kernel = FakeMHKernel(FakeSteppedKernel(step_size=0.1), log_accept_ratio=np.log(p))
  kernel = tfp.mcmc.SimpleStepSizeAdaptation(kernel, num_adaptation_steps=1, adaptation_rate=0.5)
  kr = _two_steps(kernel, torch.tensor(0.))
  expect = 0.1 * 1.5 if up else 0.1 / 1.5          # only the first step adapts
  np.testing.assert_allclose(float(kr.inner_results.accepted_results.step_size), expect, rtol=1e-6)


def test_simple_adaptation_list_step_and_chain_reduction():
  # simple_step_size_adaptation_test.py testListStep / testChainLogProbScalarTarget restated
  kernel = FakeMHKernel(FakeSteppedKernel(step_size=np.float32([0.1, 0.2, 0.3])),
                        log_accept_ratio=np.log([0.74, 0.76, 0.76]))
  kernel = tfp.mcmc.SimpleStepSizeAdaptation(kernel, num_adaptation_steps=1, adaptation_rate=1.)
  kr = _two_steps(kernel, torch.zeros(3))
  np.testing.assert_allclose(kr.inner_results.accepted_results.step_size.numpy(), [0.05, 0.4, 0.6], rtol=1e-6)
  # per-part scalar steps with a [chains] accept prob: log-mean-exp over chains
  kernel = FakeMHKernel(FakeSteppedKernel(step_size=[0.1, 0.2]), log_accept_ratio=np.log([0.5, 0.9, 0.9]))
  kernel = tfp.mcmc.SimpleStepSizeAdaptation(kernel, num_adaptation_steps=3, adaptation_rate=1.)
  kr = _two_steps(kernel, [torch.zeros(3), torch.zeros(3, 2)])
  got = kr.inner_results.accepted_results.step_size
  np.testing.assert_allclose([float(got[0]), float(got[1])], [0.2, 0.4], rtol=1e-6)   # mean 0.7667 > 0.75: doubled
  np.testing.assert_allclose([float(v) for v in kr.new_step_size], [0.4, 0.8], rtol=1e-6)


def test_simple_adaptation_finite():
  """hmc_test.py:917-952 restated on the fake kernel: every step accepts, the step size is doubled exactly
  num_adaptation_steps times and constant afterwards (the GPU version of this pin runs real HMC)."""
  n_adapt = 3
  kernel = FakeMHKernel(FakeSteppedKernel(step_size=1e-5), log_accept_ratio=0.)
  kernel = tfp.mcmc.SimpleStepSizeAdaptation(kernel, num_adaptation_steps=n_adapt, adaptation_rate=1.)
  state = torch.tensor(0.)
  kr = kernel.bootstrap_results(state)
  steps = []
  for _ in range(10):
    _, kr = kernel.one_step(state, kr)
    steps.append(float(kr.new_step_size))
  np.testing.assert_allclose(steps[n_adapt], 1e-5 * 2 ** n_adapt, atol=1e-6 * 1e-5 * 8)
  assert min(steps[n_adapt:]) == max(steps[n_adapt:])


def test_unsupported_options_raise():
  inner = FakeMHKernel(FakeSteppedKernel(step_size=1.), log_accept_ratio=0.)
  with pytest.raises(NotImplementedError):
    tfp.mcmc.DualAveragingStepSizeAdaptation(inner, 1, reduce_fn=lambda *a, **k: None)
  with pytest.raises(NotImplementedError):
    tfp.mcmc.SimpleStepSizeAdaptation(inner, 1, reduce_fn=lambda *a, **k: None)


def test_step_size_tensor_forms():
  """_engine.step_size_tensor: scalar | [D] | per-part lists | per-chain [B, 1] (also inside a list) and the
  error for per-part shapes the kernels cannot express (round-1 advisor finding)."""
  from probability_b200 import _lib
  from probability_b200.mcmc import _engine
  dev = torch.device('cpu')
  shapes = [(), (3,)]           # state parts [B] and [B, 3]: D = 4
  B, D = 5, 4
  s, kind = _engine.step_size_tensor(0.3, B, D, shapes, dev)
  assert kind == _lib.STEP_SCALAR and s.tolist() == pytest.approx([0.3])
  s, kind = _engine.step_size_tensor([0.1, np.float32([1., 2., 3.])], B, D, shapes, dev)
  assert kind == _lib.STEP_PER_DIM and s.tolist() == pytest.approx([0.1, 1., 2., 3.])
  s, kind = _engine.step_size_tensor([0.5], B, D, shapes, dev)
  assert kind == _lib.STEP_PER_DIM and s.tolist() == pytest.approx([0.5] * 4)
  pc = np.linspace(0.1, 0.5, B).astype(np.float32)
  s, kind = _engine.step_size_tensor(pc[:, None], B, D, [(4,)], dev)
  assert kind == _lib.STEP_PER_CHAIN and s.tolist() == pytest.approx(pc.tolist())
  s, kind = _engine.step_size_tensor([pc, pc[:, None]], B, D, shapes, dev)
  assert kind == _lib.STEP_PER_CHAIN and s.shape == (B,)
  with pytest.raises(ValueError):
    _engine.step_size_tensor([pc, 2 * pc[:, None]], B, D, shapes, dev)      # parts disagree
  with pytest.raises(ValueError):
    _engine.step_size_tensor([0.1, np.ones((B, 3), np.float32)], B, D, shapes, dev)   # per-chain AND per-dim
  with pytest.raises(ValueError):
    _engine.step_size_tensor([0.1, 0.2, 0.3], B, D, shapes, dev)
