"""GPU statistical parity (north_star: posterior means and variances within 4 Monte-Carlo standard
errors).  MCSE = sd / sqrt(cross-chain ESS) from the engine's own estimator, combined with the
ground truth's reported standard error where the truth is itself a Monte-Carlo estimate."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def tfp():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  import probability_b200 as tfp_
  return tfp_


def dev():
  return torch.device('cuda', 0)


def _mcse(tfp, draws):
  """draws [N, B, D] -> (mean [D], sd [D], mcse_of_mean [D], ess [D])."""
  ess = tfp.mcmc.effective_sample_size(draws, cross_chain_dims=1, filter_beyond_positive_pairs=True,
                                       filter_threshold=None).cpu().numpy().astype(np.float64)
  x = draws.double()
  mean = x.mean((0, 1)).cpu().numpy()
  sd = x.reshape(-1, x.shape[-1]).std(0).cpu().numpy()
  return mean, sd, sd / np.sqrt(ess), ess


def _sample(tfp, kernel_fn, state, warm, adapt, draws, seed):
  k0 = kernel_fn()
  k = tfp.mcmc.DualAveragingStepSizeAdaptation(k0, num_adaptation_steps=adapt)
  res = tfp.mcmc.sample_chain(1, state, kernel=k, num_burnin_steps=warm, trace_fn=None, seed=seed,
                              return_final_kernel_results=True)
  eps = float(res.final_kernel_results.new_step_size)
  st = res.all_states[0] if torch.is_tensor(res.all_states) else [s[0] for s in res.all_states]
  out = tfp.mcmc.sample_chain(draws, st, kernel=k0.copy(step_size=eps), seed=seed + 1,
                              trace_fn=lambda _, kr: (kr.log_accept_ratio, kr.is_accepted))
  return out, eps


@pytest.mark.parametrize('sampler', ['nuts', 'hmc'])
def test_eight_schools_posterior_vs_stan_ground_truth(tfp, sampler):
  gt = json.load(open(os.path.join(G, 'eight_schools_ground_truth.json')))
  tg = tfp.targets.EightSchools()
  B = 2048
  state = [torch.zeros(B, device=dev()), torch.zeros(B, device=dev()), torch.ones(B, 8, device=dev())]
  if sampler == 'nuts':
    mk = lambda: tfp.mcmc.NoUTurnSampler(tg, step_size=0.2, max_tree_depth=8)
  else:
    mk = lambda: tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.2, num_leapfrog_steps=8)
  out, eps = _sample(tfp, mk, state, warm=400, adapt=320, draws=250, seed=3)
  mu, tau, z = out.all_states
  theta = mu[..., None] + torch.exp(tau)[..., None] * z      # school_effects = mu + e^tau z
  draws = torch.cat([mu[..., None], tau[..., None], theta], -1).contiguous()
  mean, sd, mcse, ess = _mcse(tfp, draws)
  truth = np.concatenate([[gt['IDENTITY_AVG_EFFECT_MEAN']], [gt['IDENTITY_LOG_STDDEV_MEAN']],
                          gt['IDENTITY_SCHOOL_EFFECTS_MEAN']])
  se = np.concatenate([[gt['IDENTITY_AVG_EFFECT_MEAN_STANDARD_ERROR']],
                       [gt['IDENTITY_LOG_STDDEV_MEAN_STANDARD_ERROR']],
                       gt['IDENTITY_SCHOOL_EFFECTS_MEAN_STANDARD_ERROR']])
  true_sd = np.concatenate([[gt['IDENTITY_AVG_EFFECT_STANDARD_DEVIATION']],
                            [gt['IDENTITY_LOG_STDDEV_STANDARD_DEVIATION']],
                            gt['IDENTITY_SCHOOL_EFFECTS_STANDARD_DEVIATION']])
  assert ess.min() > 2000, ess
  zscore = np.abs(mean - truth) / np.sqrt(mcse ** 2 + se ** 2)
  assert zscore.max() < 4.0, (zscore, mean, truth)
  # variances: sd within 4 standard errors of the sd estimate (~ sd / sqrt(2 ESS)) plus the truth's own noise
  sd_se = np.sqrt((sd / np.sqrt(2 * ess)) ** 2 + (true_sd / np.sqrt(2 * 20000 * 0.25)) ** 2)
  assert (np.abs(sd - true_sd) / sd_se).max() < 4.0, (sd, true_sd)
  acc = torch.exp(torch.minimum(out.trace[0], torch.zeros_like(out.trace[0]))).mean().item()
  assert 0.6 < acc < 0.95, acc


def test_dense_gaussian_posterior_moments(tfp):
  """Ill-conditioned 100-d Gaussian (C2): mean 0 and marginal variances diag(Sigma) within 4 MCSE."""
  tg = tfp.targets.IllConditionedGaussian()
  B = 1024
  rng = np.random.default_rng(5)
  L = np.linalg.cholesky(tg.covariance)
  x0 = torch.tensor((rng.standard_normal((B, 100)) @ L.T).astype(np.float32), device=dev())
  out, eps = _sample(tfp, lambda: tfp.mcmc.NoUTurnSampler(tg, step_size=0.158, max_tree_depth=10), x0,
                     warm=120, adapt=100, draws=200, seed=9)
  draws = out.all_states
  mean, sd, mcse, ess = _mcse(tfp, draws)
  true_sd = np.sqrt(np.diag(tg.covariance))
  assert (np.abs(mean) / mcse).max() < 4.5, (np.abs(mean) / mcse).max()      # 100 dims: allow the max of 100 z's
  sd_se = sd / np.sqrt(2 * np.minimum(ess, draws.shape[0] * B))
  # second moments mix differently from means; use the ESS of x^2
  ess2 = tfp.mcmc.effective_sample_size((draws ** 2).contiguous(), cross_chain_dims=1,
                                        filter_beyond_positive_pairs=True, filter_threshold=None).cpu().numpy()
  var = sd ** 2
  var_se = np.sqrt(2.0) * true_sd ** 2 / np.sqrt(ess2)
  assert (np.abs(var - true_sd ** 2) / var_se).max() < 4.5, (np.abs(var - true_sd ** 2) / var_se).max()
  assert 0.5 < eps < 1.2


def test_logistic_nuts_and_hmc_agree(tfp):
  """Same posterior through two different transition kernels (NUTS vs HMC), 1000x25 logistic (C3 shape)."""
  X, y = tfp.targets.synthetic_logistic_data(1000, 24, seed=0)
  tg = tfp.targets.LogisticRegression(X, y)
  B = 1024
  st = torch.zeros(B, 25, device=dev())
  a, _ = _sample(tfp, lambda: tfp.mcmc.NoUTurnSampler(tg, step_size=0.1, max_tree_depth=8), st, 200, 160, 150, 1)
  b, _ = _sample(tfp, lambda: tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.1, num_leapfrog_steps=10), st, 200, 160,
                 150, 7)
  ma, sa, ea, _ = _mcse(tfp, a.all_states)
  mb, sb, eb, _ = _mcse(tfp, b.all_states)
  z = np.abs(ma - mb) / np.sqrt(ea ** 2 + eb ** 2)
  assert z.max() < 4.0, z
  np.testing.assert_allclose(sa, sb, rtol=0.05)
  rhat = tfp.mcmc.potential_scale_reduction(a.all_states, split_chains=True).cpu().numpy()
  assert rhat.max() < 1.02


def test_stochastic_volatility_sp500_vs_stan_ground_truth(tfp):
  """C4 on the reference's embedded S&P 500 series: the three scalar parameters and a sample of the
  log-volatility path against the Stan ground truth (gym ground_truth/stochastic_volatility_sp500.py)."""
  z = np.load(os.path.join(G, 'sv_sp500.npz'))
  tg = tfp.targets.StochasticVolatility(z['centered_returns'])
  B = 296
  u0 = torch.zeros(B, tg.dim, device=dev())
  u0[:, 0] = 3.0                                           # phi ~ 0.9
  u0[:, 1] = float(np.log(np.var(z['centered_returns'])))
  u0[:, 2] = -1.0
  out, eps = _sample(tfp, lambda: tfp.mcmc.NoUTurnSampler(tg, step_size=0.02, max_tree_depth=8), u0,
                     warm=300, adapt=250, draws=120, seed=21)
  u = out.all_states
  c = tg.constrain(u)
  phi, m, s = c[..., 0], c[..., 1], c[..., 2]
  scal = torch.stack([phi, m, s], -1).contiguous()
  mean, sd, mcse, ess = _mcse(tfp, scal)
  truth = np.array([float(z['identity_persistence_of_volatility_mean']), float(z['identity_mean_log_volatility_mean']),
                    float(z['identity_white_noise_shock_scale_mean'])])
  se = np.array([float(z['identity_persistence_of_volatility_mean_standard_error']),
                 float(z['identity_mean_log_volatility_mean_standard_error']),
                 float(z['identity_white_noise_shock_scale_mean_standard_error'])])
  zs = np.abs(mean - truth) / np.sqrt(mcse ** 2 + se ** 2)
  assert zs.max() < 4.0, (zs, mean, truth, ess)
  # log-volatility path h_t + m at a few time points (recomputed with torch from the draws)
  zz = c[..., 3:]
  h = torch.empty_like(zz)
  h[..., 0] = s * zz[..., 0] / torch.sqrt(1 - phi * phi)
  for t in range(1, 40):
    h[..., t] = phi * h[..., t - 1] + s * zz[..., t]
  lv = (h[..., :40] + m[..., None])[..., ::8].contiguous()
  mean, sd, mcse, ess = _mcse(tfp, lv)
  truth = z['identity_log_volatility_mean'][:40:8]
  se = z['identity_log_volatility_mean_standard_error'][:40:8]
  zs = np.abs(mean - truth) / np.sqrt(mcse ** 2 + se ** 2)
  assert zs.max() < 4.0, (zs, mean, truth)
