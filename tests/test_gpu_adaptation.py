"""GPU tests of the step-size adaptation kernels wrapped around the CUDA transitions (through the C ABI):
SimpleStepSizeAdaptation against the reference's finite-adaptation pin (tfp/mcmc/hmc_test.py:917-952) and the
oracle; DualAveragingStepSizeAdaptation with per-chain `[chains, 1]` and per-part step sizes
(dual_averaging_step_size_adaptation.py:353-475) on the real kernels."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu

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


def test_simple_adaptation_finite_pin(tfp):
  """hmc_test.py:917-952: HMC on N(0, 1) with a tiny initial step size accepts every proposal, so the step size is
  multiplied by (1 + adaptation_rate) = 2 exactly num_adaptation_steps times and then stays constant."""
  num_results, n_adapt, eps0 = 10, 3, 1e-5
  tg = tfp.targets.DenseGaussian(covariance=np.ones((1, 1)))
  k = tfp.mcmc.SimpleStepSizeAdaptation(
      tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=eps0, num_leapfrog_steps=2), num_adaptation_steps=n_adapt,
      adaptation_rate=1.)
  res = tfp.mcmc.sample_chain(num_results, torch.zeros(1, 1, device=dev()), kernel=k, num_burnin_steps=0,
                              trace_fn=lambda _, kr: kr.new_step_size, seed=7)
  steps = res.trace.cpu().numpy().reshape(-1)
  np.testing.assert_allclose(steps[n_adapt], eps0 * 2 ** n_adapt, atol=1e-6 * eps0)
  np.testing.assert_allclose(steps[:n_adapt], eps0 * 2 ** np.arange(1, n_adapt + 1), rtol=1e-6)
  assert steps[n_adapt:].min() == steps[n_adapt:].max()


def test_simple_adaptation_matches_oracle(tfp):
  tg = tfp.targets.EightSchools()
  og = otargets.EightSchools()
  rng = np.random.default_rng(3)
  x = (np.array([0, 0] + [1] * 8) + 0.5 * rng.standard_normal((64, 10))).astype(np.float32)
  state = [torch.tensor(x[:, 0], device=dev()), torch.tensor(x[:, 1], device=dev()), torch.tensor(x[:, 2:], device=dev())]
  k = tfp.mcmc.SimpleStepSizeAdaptation(
      tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.2, num_leapfrog_steps=3), num_adaptation_steps=12,
      adaptation_rate=0.05)
  res = tfp.mcmc.sample_chain(16, state, kernel=k, trace_fn=lambda _, kr: kr.new_step_size, seed=5)
  ad = omcmc.SimpleAdaptation(0.2, 12, adaptation_rate=0.05)
  _, ref_trace, _ = omcmc.sample_chain(og, 'hmc', x, 16, step_size=0.2, num_leapfrog_steps=3, seed=5,
                                       dual_averaging=ad)
  # trace entry r of the oracle is the step USED by transition r; new_step_size[r] is the one for r + 1
  ref_used = np.array([r['step_size'] for r in ref_trace], np.float32)
  got = res.trace.cpu().numpy().reshape(-1)
  np.testing.assert_allclose(got[:-1], ref_used[1:], rtol=1e-5)
  assert got[12:].min() == got[12:].max()
  assert len(set(np.round(np.log(got[:12] / 0.2) / np.log(1.05)).astype(int))) > 1   # it went up and down


def test_dual_averaging_per_chain_step_sizes_adapt_independently(tfp):
  """A `[chains, 1]` step size makes every chain adapt on its own accept ratio (num_reduce_dims = 0,
  dual_averaging...py:419-435): the run equals single-chain runs of the same chains (global chain index RNG)."""
  cov = np.array([[1.0, 0.5, 0.1], [0.5, 2.0, -0.3], [0.1, -0.3, 0.5]])
  tg = tfp.targets.DenseGaussian(covariance=cov)
  B = 6
  rng = np.random.default_rng(0)
  x = torch.tensor(rng.standard_normal((B, 3)).astype(np.float32), device=dev())
  eps0 = np.linspace(0.05, 1.5, B).astype(np.float32)[:, None]

  def run(xs, eps, shard):
    k = tfp.mcmc.DualAveragingStepSizeAdaptation(
        tfp.mcmc.NoUTurnSampler(tg, step_size=torch.tensor(eps, device=dev()), max_tree_depth=5,
                                experimental_chain_shard=shard), num_adaptation_steps=15)
    return tfp.mcmc.sample_chain(20, xs, kernel=k, seed=11,
                                 trace_fn=lambda _, kr: (kr.new_step_size, kr.inner_results.leapfrogs_taken))

  full = run(x, eps0, None)
  steps = full.trace[0].cpu().numpy()
  assert steps.shape == (20, B, 1)
  assert (np.abs(np.log(steps[14, :, 0] / steps[14, 0, 0])) < 3.0).all()      # all chains found a similar scale
  assert len(np.unique(steps[5])) == B                                        # ... on their own paths
  np.testing.assert_array_equal(steps[15:], np.broadcast_to(steps[14], steps[15:].shape))   # frozen
  for c in (0, 3):   # two-chain shards (a one-chain shard would take the scalar device path: other rounding)
    two = run(x[c:c + 2], eps0[c:c + 2], tfp.mcmc.ChainShard(c, B))
    np.testing.assert_array_equal(two.all_states.cpu().numpy(), full.all_states.cpu().numpy()[:, c:c + 2])
    np.testing.assert_array_equal(two.trace[0].cpu().numpy(), steps[:, c:c + 2])
  # first update against the closed form (:437-452): eps_1 = exp(log(10 eps_0) - (0.75 - p) / ((10 + 1) 0.05))
  k = tfp.mcmc.DualAveragingStepSizeAdaptation(
      tfp.mcmc.NoUTurnSampler(tg, step_size=torch.tensor(eps0, device=dev()), max_tree_depth=5),
      num_adaptation_steps=15)
  kr = k.bootstrap_results(x)
  _, kr1 = k.one_step(x, kr, seed=orng.key(2))
  p = np.exp(np.minimum(kr1.inner_results.log_accept_ratio.cpu().numpy(), 0.))
  expect = np.exp(np.log(10. * eps0[:, 0]) - (0.75 - p) / (11. * 0.05))
  np.testing.assert_allclose(kr1.new_step_size.cpu().numpy()[:, 0], expect, rtol=2e-5)


def test_dual_averaging_per_part_step_sizes_keep_their_ratio(tfp):
  """hmc_test.py:880-915 (list step sizes): every part is scaled by the same chain-reduced accept statistic, so the
  ratio between the parts' step sizes stays what it was; the fused scalar run gives the same accept statistics."""
  tg = tfp.targets.EightSchools()
  rng = np.random.default_rng(1)
  x = (np.array([0, 0] + [1] * 8) + 0.3 * rng.standard_normal((128, 10))).astype(np.float32)
  state = [torch.tensor(x[:, 0], device=dev()), torch.tensor(x[:, 1], device=dev()), torch.tensor(x[:, 2:], device=dev())]
  k = tfp.mcmc.DualAveragingStepSizeAdaptation(
      tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=[0.1, 0.2, 0.05], num_leapfrog_steps=3), num_adaptation_steps=10)
  res = tfp.mcmc.sample_chain(12, state, kernel=k, trace_fn=lambda _, kr: kr.new_step_size, seed=3)
  s = [v.cpu().numpy() for v in res.trace]
  assert len(s) == 3 and s[0].shape == (12,)
  np.testing.assert_allclose(s[1] / s[0], 2.0, rtol=1e-5)
  np.testing.assert_allclose(s[2] / s[0], 0.5, rtol=1e-5)
  assert s[0][11] == s[0][10] and s[0][3] != s[0][2]
