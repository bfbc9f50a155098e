"""Config-scale parity (BASELINE.json configs at their real shapes, bounded chain counts so the CPU oracle finishes in
seconds): the kernels that the bench times are compared with the oracle at the depth / series length / row count of
the named configs, not only at toy sizes.

  C2  asynchronous-lane tcgen05 NUTS, 100-d ill-conditioned Gaussian, max_tree_depth = 10  -> oracle directly
  C3  logistic regression 1000 x 25, NUTS depth 10 + DualAveraging, a few transitions      -> oracle seed chain
  C4  stochastic volatility T = 2516, NUTS depth 10                                         -> oracle
  C5  row-sharded logistic gradient at 125,000 rows x 100 weights (one rank's shard)        -> float64
Measured agreement is printed (pytest -s) and pinned close to it.
"""
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


def test_c2_async_tile_nuts_depth10_matches_oracle(tfp):
  """The bench's kernel (tile_nuts_async_kernel, depth 10, adapted step size 0.74) against the oracle itself -- not
  only transitively through the lock-step tile kernel."""
  tg = tfp.targets.IllConditionedGaussian()
  og = otargets.DenseGaussian(tg.precision, tg.log_normalizer)
  B = 320
  rng = np.random.default_rng(5)
  x = (rng.standard_normal((B, 100)) @ np.linalg.cholesky(tg.covariance).T).astype(np.float32)
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=10)
  xt = torch.tensor(x, device=dev())
  seed = orng.key(101)
  st, kr = k.one_step(xt, k.bootstrap_results(xt), seed=seed)
  lp0, g0 = og.logp_grad(x)
  ref = omcmc.nuts_one_step(og, x, lp0, g0, 0.74, seed, max_tree_depth=10)
  nl = kr.leapfrogs_taken.cpu().numpy()
  same = nl == ref['leapfrogs_taken']
  print('C2 depth-10 async tile vs oracle: identical trees %.4f (mean leapfrogs %.1f, max %d)' % (
      same.mean(), nl.mean(), nl.max()))
  assert nl.max() >= 511                     # the deep trees of this config are exercised
  assert same.mean() >= 0.97
  close = np.isclose(st.cpu().numpy(), ref['state'], rtol=5e-3, atol=5e-2).all(1)
  assert close[same].mean() >= 0.97
  for f in ('is_accepted', 'reach_max_depth', 'has_divergence'):
    assert (getattr(kr, f).cpu().numpy() == ref[f])[same & close].all(), f


def test_c3_logistic_nuts_depth10_dual_averaging_matches_oracle(tfp):
  X, y = otargets.synthetic_logistic_data(1000, 24, seed=0)
  tg = tfp.targets.LogisticRegression(X[:, :-1], y)
  og = otargets.LogisticRegression(X, y)
  B, n = 192, 4
  x = np.zeros((B, 25), np.float32)
  k = tfp.mcmc.DualAveragingStepSizeAdaptation(
      tfp.mcmc.NoUTurnSampler(tg, step_size=0.1, max_tree_depth=10), num_adaptation_steps=3)
  res = tfp.mcmc.sample_chain(n, torch.tensor(x, device=dev()), kernel=k, seed=23,
                              trace_fn=lambda _, kr: (kr.inner_results.step_size, kr.inner_results.leapfrogs_taken))
  da = omcmc.DualAveraging(0.1, 3)
  states, trace, _ = omcmc.sample_chain(og, 'nuts', x, n, step_size=0.1, max_tree_depth=10, seed=23, dual_averaging=da)
  steps = res.trace[0].cpu().numpy()
  ref_steps = np.array([r['step_size'] for r in trace], np.float32)
  nl = res.trace[1].cpu().numpy()
  ref_nl = np.stack([r['leapfrogs_taken'] for r in trace])
  same0 = (nl[0] == ref_nl[0]).mean()
  print('C3 depth-10 + DA: identical trees in transition 0: %.4f; step sizes %s vs oracle %s; max leapfrogs %d' % (
      same0, steps, ref_steps, nl.max()))
  assert same0 >= 0.97
  np.testing.assert_allclose(steps[:2], ref_steps[:2], rtol=1e-3)      # step 1 depends on transition 0's accept stat
  np.testing.assert_allclose(steps, ref_steps, rtol=0.05)
  # chains whose whole tree history matches end in the same state
  hist = (nl == ref_nl).all(0)
  assert hist.mean() >= 0.85
  got = res.all_states.cpu().numpy()[-1]
  close = np.isclose(got, states[-1], rtol=5e-3, atol=5e-3).all(1)
  assert close[hist].mean() >= 0.97


def test_c4_stochastic_volatility_T2516_nuts_matches_oracle(tfp):
  yv = tfp.targets.synthetic_sv_returns(T=2516, seed=0)
  tg = tfp.targets.StochasticVolatility(yv)
  og = otargets.StochasticVolatility(yv)
  B = 12
  rng = np.random.default_rng(4)
  x = (0.05 * rng.standard_normal((B, 2519))).astype(np.float32)
  x[:, 0] += 2.0
  x[:, 1] += 5.0
  # warm up on the GPU (untimed, unchecked) so that the compared transition grows the ~100-leaf trees of the config
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.05, max_tree_depth=10)
  def parts(a):   # the gym model's four state parts: one momentum key per part (hmc.py:684-695)
    a = torch.tensor(a, device=dev())
    return [a[:, 0].contiguous(), a[:, 1].contiguous(), a[:, 2].contiguous(), a[:, 3:].contiguous()]

  warm = tfp.mcmc.sample_chain(1, parts(x), num_burnin_steps=70, seed=3, trace_fn=None,
                               kernel=tfp.mcmc.DualAveragingStepSizeAdaptation(k, num_adaptation_steps=60),
                               return_final_kernel_results=True)
  eps = float(warm.final_kernel_results.new_step_size)
  x = np.ascontiguousarray(torch.cat([s[0].reshape(B, -1) for s in warm.all_states], 1).cpu().numpy())
  k = k.copy(step_size=eps)
  seed = orng.key(77)
  st, kr = k.one_step(parts(x), k.bootstrap_results(parts(x)), seed=seed)
  st = torch.cat([s.reshape(B, -1) for s in st], 1)
  lp0, g0 = og.logp_grad(x)
  ref = omcmc.nuts_one_step(og, x, lp0, g0, np.float32(eps), seed, max_tree_depth=10)
  nl = kr.leapfrogs_taken.cpu().numpy()
  same = nl == ref['leapfrogs_taken']
  print('C4 T=2516 depth-10 vs oracle (eps %.4f): identical trees %d of %d, leapfrogs %s' % (eps, same.sum(), B, nl))
  assert nl.max() >= 63
  assert same.mean() >= 0.9      # measured 12 of 12 (round 2)
  got = st.cpu().numpy()
  rel = np.linalg.norm(got - ref['state'], axis=1) / np.linalg.norm(ref['state'], axis=1)
  print('   relative L2 state error per chain %s' % np.array2string(rel, precision=2))
  # float32 rounding grows along a ~100-leapfrog trajectory of a 2519-d system: same tree => same multinomial pick,
  # state equal up to the accumulated rounding
  assert (rel[same] < 1e-4).all()   # measured 2e-7 .. 8e-7
  np.testing.assert_allclose(kr.energy.cpu().numpy()[same], ref['energy'][same], rtol=1e-3)


def test_c5_rowshard_gradient_at_shard_size_matches_float64(tfp):
  """One rank's share of C5: 125,000 rows x 100 weights, tcgen05 path (B >= 128), against float64."""
  n, d, B = 125000, 99, 256
  rng = np.random.default_rng(11)
  X = rng.standard_normal((n, d)).astype(np.float32)
  theta_true = (0.1 * rng.standard_normal(d + 1)).astype(np.float32)
  y = (rng.random(n) < 1 / (1 + np.exp(-(X @ theta_true[:-1] + theta_true[-1])))).astype(np.float32)
  tg = tfp.targets.RowShardedLogisticRegression(X, y)
  th = (theta_true + 2e-3 * rng.standard_normal((B, d + 1))).astype(np.float32)
  lp, g = tg.log_prob_and_grad(torch.tensor(th, device=dev()))
  Xb = np.concatenate([X, np.ones((n, 1), np.float32)], 1).astype(np.float64)
  z = th.astype(np.float64) @ Xb.T                                    # [B, n]
  ll = (y[None] * z - np.logaddexp(0, z)).sum(1)
  lp64 = ll - 0.5 * (th.astype(np.float64) ** 2).sum(1) - 0.5 * (d + 1) * np.log(2 * np.pi)
  g64 = (y[None] - 1 / (1 + np.exp(-z))) @ Xb - th
  err_lp = np.max(np.abs(lp.cpu().numpy() - lp64) / np.abs(lp64))
  scale = np.abs(g64).max(axis=1, keepdims=True)
  err_g = np.max(np.abs(g.cpu().numpy() - g64) / scale)
  print('C5 shard 125000 x 100, %d chains: rel err logp %.2e, grad %.2e' % (B, err_lp, err_g))
  assert err_lp < 5e-6
  assert err_g < 5e-5
  # deterministic partial reduction: identical bits on a second launch (what keeps the replicas in lock-step)
  lp2, g2 = tg.log_prob_and_grad(torch.tensor(th, device=dev()))
  assert torch.equal(g, g2) and torch.equal(lp, lp2)
