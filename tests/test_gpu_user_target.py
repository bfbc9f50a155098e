"""User-defined targets on the GPU (the reference's arbitrary `target_log_prob_fn`, tfp/mcmc/hmc.py:413-415;
inference_gym model_contract.md): CUDA source for one chain's log-prob + gradient, compiled at run time into the
chain kernels.  Checked against the oracle: Eight Schools re-written by a "user" gives the named target's numbers
(log-prob / gradient to float32 rounding, the same HMC decisions and NUTS trees), a dense Gaussian handed over through
`data` samples the right moments under sample_chain + dual averaging, multi-element-per-lane (D = 40) and cooperative
(all 32 lanes) sources work."""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu

from oracle import mcmc as omcmc  # noqa: E402
from oracle import rng as orng  # noqa: E402
from oracle import targets as otargets  # noqa: E402

EIGHT_SCHOOLS_SRC = r'''
// Eight Schools, non-centred (tfp/mcmc/eight_schools_hmc.py:41-57): x = [mu, tau, z_0..z_{J-1}], data = [y | sigma]
__device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data) {
  const int J = n_data / 2;
  const float* y = data;
  const float* sigma = data + J;
  const float HL2P = 0.91893853320467274178f;
  const float mu = x[0], tau = x[1], e = expf(tau);
  float lp = (-0.5f * (mu / 10.f) * (mu / 10.f) - (HL2P + logf(10.f))) + (-0.5f * (tau - 5.f) * (tau - 5.f) - HL2P);
  float gmu = -mu / 100.f, gtau = -(tau - 5.f);
  for (int i = 0; i < J; ++i) {
    const float z = x[2 + i];
    const float r = (y[i] - (mu + e * z)) / sigma[i];
    const float w = r / sigma[i];
    lp += (-0.5f * z * z - HL2P) + (-0.5f * r * r - (HL2P + logf(sigma[i])));
    gmu += w;
    gtau += e * w * z;
    g[2 + i] = -z + e * w;
  }
  g[0] = gmu;
  g[1] = gtau;
  return lp;
}
'''

DENSE_SRC = r'''
// N(loc, P^-1): data = [D | loc[D] | P[D*D]]
__device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data) {
  const int D = (int)data[0];
  const float* loc = data + 1;
  const float* P = data + 1 + D;
  float q = 0.f;
  for (int i = 0; i < D; ++i) {
    float s = 0.f;
    for (int j = 0; j < D; ++j) s = fmaf(P[i * D + j], x[j] - loc[j], s);
    g[i] = -s;
    q = fmaf(x[i] - loc[i], s, q);
  }
  return -0.5f * q;
}
'''

COOP_SRC = r'''
// independent normals with per-dimension scale data[d], evaluated by all 32 lanes (lane-strided)
__device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data, int lane) {
  float lp = 0.f;
  for (int d = lane; d < n_data; d += 32) {
    const float z = x[d] / data[d];
    g[d] = -z / data[d];
    lp -= 0.5f * z * z;
  }
  return pb2::warp_sum(lp);
}
'''


@pytest.fixture(scope='module')
def tfp():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  import probability_b200 as tfp_
  return tfp_


def dev():
  return torch.device('cuda', 0)


def _schools(tfp):
  es = tfp.targets.EightSchools
  data = np.concatenate([np.asarray(es.TREATMENT_EFFECTS, np.float32), np.asarray(es.TREATMENT_STDDEVS, np.float32)])
  return tfp.targets.UserTarget(10, EIGHT_SCHOOLS_SRC, data=data, part_sizes=[1, 1, 8])


def _parts(x):
  xt = torch.tensor(x, device=dev())
  return [xt[:, 0].contiguous(), xt[:, 1].contiguous(), xt[:, 2:].contiguous()]


def _flat(state):
  return torch.cat([s.reshape(s.shape[0], -1) for s in state], 1).cpu().numpy()


def test_user_eight_schools_logp_grad_matches_oracle(tfp):
  tg = _schools(tfp)
  rng = np.random.default_rng(0)
  x = (np.array([0, 0] + [1] * 8) + 0.5 * rng.standard_normal((257, 10))).astype(np.float32)
  lp, g = tg.log_prob_and_grad(torch.tensor(x, device=dev()))
  lp64, g64 = otargets.EightSchools(dtype=np.float64).logp_grad(x.astype(np.float64))
  lp32, g32 = otargets.EightSchools().logp_grad(x)
  rel = lambda a, b: np.max(np.abs(a - b) / (np.abs(b) + 1e-3 * np.max(np.abs(b))))
  assert rel(lp.cpu().numpy(), lp64) < max(5 * rel(lp32, lp64), 1e-6)
  assert rel(g.cpu().numpy(), g64) < max(5 * rel(g32, g64), 1e-5)
  # and the named target, through the same entry point
  lpn, gn = tfp.targets.EightSchools().log_prob_and_grad(torch.tensor(x, device=dev()))
  np.testing.assert_allclose(lp.cpu().numpy(), lpn.cpu().numpy(), rtol=2e-5, atol=2e-5)
  np.testing.assert_allclose(g.cpu().numpy(), gn.cpu().numpy(), rtol=2e-4, atol=2e-4)


def test_user_eight_schools_transitions_match_oracle(tfp):
  tg = _schools(tfp)
  es = otargets.EightSchools()
  rng = np.random.default_rng(1)
  x = (np.array([0, 0] + [1] * 8) + 0.5 * rng.standard_normal((128, 10))).astype(np.float32)
  lp0, g0 = es.logp_grad(x)
  state = _parts(x)
  # HMC + Metropolis-Hastings
  k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.4, num_leapfrog_steps=3)
  seed = orng.key(11)
  new_state, kr = k.one_step(state, k.bootstrap_results(state), seed=seed)
  ref = omcmc.hmc_one_step(es, x, lp0, g0, 0.4, 3, seed)
  acc = kr.is_accepted.cpu().numpy()
  agree = acc == ref['is_accepted']
  assert agree.mean() >= 0.97
  np.testing.assert_allclose(_flat(kr.proposed_state), ref['proposed_state'], rtol=1e-4, atol=1e-4)
  np.testing.assert_allclose(_flat(new_state)[agree], ref['state'][agree], rtol=1e-4, atol=1e-4)
  # NUTS
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.3, max_tree_depth=6)
  seed = orng.key(31)
  new_state, kr = k.one_step(state, k.bootstrap_results(state), seed=seed)
  ref = omcmc.nuts_one_step(es, x, lp0, g0, 0.3, seed, max_tree_depth=6)
  same = kr.leapfrogs_taken.cpu().numpy() == ref['leapfrogs_taken']
  print('user Eight Schools NUTS: identical trees %.4f' % same.mean())
  assert same.mean() >= 0.97
  np.testing.assert_allclose(_flat(new_state)[same], ref['state'][same], rtol=2e-3, atol=2e-3)


def test_user_dense_gaussian_posterior_under_adaptive_nuts(tfp):
  """sample_chain + DualAveragingStepSizeAdaptation on a user target: moments of a 3-d correlated normal
  (the shape of nuts_test.py:310-331) within 5 standard errors."""
  cov = np.array([[1.0, 0.5, 0.1], [0.5, 2.0, -0.3], [0.1, -0.3, 0.5]])
  loc = np.array([1.0, -2.0, 0.5])
  P = np.linalg.inv(cov)
  data = np.concatenate([[3.0], loc, P.reshape(-1)]).astype(np.float32)
  tg = tfp.targets.UserTarget(3, DENSE_SRC, data=data)
  B, R = 512, 200
  x0 = torch.zeros(B, 3, device=dev())
  k = tfp.mcmc.DualAveragingStepSizeAdaptation(tfp.mcmc.NoUTurnSampler(tg, step_size=0.5, max_tree_depth=6),
                                               num_adaptation_steps=80)
  res = tfp.mcmc.sample_chain(R, x0, kernel=k, num_burnin_steps=100, seed=5, trace_fn=None)
  s = res.cpu().numpy().reshape(-1, 3).astype(np.float64)   # trace_fn=None: sample_chain returns the states
  n_eff = B * R / 4.0
  assert np.all(np.abs(s.mean(0) - loc) < 5 * np.sqrt(np.diag(cov) / n_eff))
  np.testing.assert_allclose(np.cov(s.T), cov, atol=0.08)
  # the oracle's dense Gaussian gives the same log-prob / gradient up to the normaliser
  xs = np.random.default_rng(3).standard_normal((64, 3)).astype(np.float32)
  lp, g = tg.log_prob_and_grad(torch.tensor(xs, device=dev()))
  o = otargets.DenseGaussian(P.astype(np.float32), 0.0, loc.astype(np.float32))
  lpo, go = o.logp_grad(xs)
  np.testing.assert_allclose(lp.cpu().numpy(), lpo, rtol=2e-5, atol=2e-5)
  np.testing.assert_allclose(g.cpu().numpy(), go, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('D', [40, 100, 256])
def test_user_cooperative_source_and_several_elements_per_lane(tfp, D):
  scale = (0.5 + np.arange(D) / D).astype(np.float32)
  tg = tfp.targets.UserTarget(D, COOP_SRC, data=scale, cooperative=True)
  x = np.random.default_rng(D).standard_normal((70, D)).astype(np.float32)
  lp, g = tg.log_prob_and_grad(torch.tensor(x, device=dev()))
  z = x.astype(np.float64) / scale
  np.testing.assert_allclose(lp.cpu().numpy(), -0.5 * (z ** 2).sum(1), rtol=2e-5)
  np.testing.assert_allclose(g.cpu().numpy(), -z / scale, rtol=2e-6, atol=1e-6)
  # fused HMC run == the step loop on the user target (same kernels either way)
  k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.2, num_leapfrog_steps=4)
  xt = torch.tensor(x, device=dev())
  fused = tfp.mcmc.sample_chain(3, xt, kernel=k, trace_fn=lambda _, kr: kr.is_accepted, seed=2)
  loop = tfp.mcmc.sample_chain(3, xt, kernel=k, trace_fn=lambda _, kr: kr.is_accepted | False, seed=2)
  np.testing.assert_array_equal(fused.all_states.cpu().numpy(), loop.all_states.cpu().numpy())
  assert fused.trace.float().mean() > 0.6


def test_user_target_preconditioned_nuts_matches_oracle(tfp):
  """Diagonal preconditioning wraps the user's function like any named target (ScaledT, third run-time build):
  PreconditionedNoUTurnSampler on the user-written Eight Schools == the oracle with that mass matrix."""
  tg = _schools(tfp)
  es = otargets.EightSchools()
  rng = np.random.default_rng(2)
  x = (np.array([0, 0] + [1] * 8) + 0.5 * rng.standard_normal((96, 10))).astype(np.float32)
  var = (0.5 + rng.random(10)).astype(np.float32)
  t = lambda a: torch.tensor(np.asarray(a), device=dev())
  md = tfp.experimental.mcmc.DiagonalMomentum([t(var[0:1].reshape(())), t(var[1:2].reshape(())), t(var[2:])])
  k = tfp.experimental.mcmc.PreconditionedNoUTurnSampler(tg, step_size=0.25, max_tree_depth=6, momentum_distribution=md)
  state = _parts(x)
  seed = orng.key(29)
  new_state, kr = k.one_step(state, k.bootstrap_results(state), seed=seed)
  lp0, g0 = es.logp_grad(x)
  ref = omcmc.nuts_one_step(es, x, lp0, g0, 0.25, seed, max_tree_depth=6, inv_mass=var)
  same = kr.leapfrogs_taken.cpu().numpy() == ref['leapfrogs_taken']
  print('user Eight Schools, preconditioned NUTS: identical trees %.4f' % same.mean())
  assert same.mean() >= 0.97
  close = np.isclose(_flat(new_state), ref['state'], rtol=3e-3, atol=3e-3).all(axis=1)
  assert close[same].mean() >= 0.97


GAMMA_SRC = r'''
// independent Gamma(shape = data[2 d], rate = data[2 d + 1]) on x_d > 0, written in its own (constrained) coordinates
__device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data) {
  float lp = 0.f;
  for (int d = 0; d < n_data / 2; ++d) {
    const float a = data[2 * d], b = data[2 * d + 1];
    lp += (a - 1.f) * logf(x[d]) - b * x[d];
    g[d] = (a - 1.f) / x[d] - b;
  }
  return lp;
}
'''


def test_user_target_behind_bijectors(tfp):
  """TransformedTransitionKernel(Exp) over a user target written on x > 0 (TransformedT, second run-time build): the
  log-prob / gradient in the unconstrained space equal the closed form with the Jacobian, and adaptive NUTS samples
  the Gamma moments."""
  ab = np.array([[3.0, 2.0], [5.0, 1.0]], np.float32)
  tg = tfp.targets.UserTarget(2, GAMMA_SRC, data=ab.reshape(-1))
  bij = tfp.bijectors.Exp()
  B = 1024
  x0 = torch.ones(B, 2, device=dev())
  k = tfp.mcmc.TransformedTransitionKernel(tfp.mcmc.NoUTurnSampler(tg, step_size=0.3, max_tree_depth=6), bij)
  kr = k.bootstrap_results(x0)
  # at u = 0: lp = sum_d -b_d + fldj (= u = 0); grad_u = (a - 1) - b e^u + 1 = a - b
  np.testing.assert_allclose(kr.inner_results.target_log_prob.cpu().numpy(), -ab[:, 1].sum() * np.ones(B), rtol=1e-6)
  g = kr.inner_results.grads_target_log_prob
  g = g[0] if isinstance(g, (list, tuple)) else g
  np.testing.assert_allclose(g.cpu().numpy(), np.tile(ab[:, 0] - ab[:, 1], (B, 1)), rtol=1e-6, atol=1e-6)
  ak = tfp.mcmc.DualAveragingStepSizeAdaptation(k, num_adaptation_steps=80)
  res = tfp.mcmc.sample_chain(150, x0, kernel=ak, num_burnin_steps=100, seed=9, trace_fn=None)
  s = res.cpu().numpy().reshape(-1, 2).astype(np.float64)
  assert (s > 0).all()
  mean, var = ab[:, 0] / ab[:, 1], ab[:, 0] / ab[:, 1] ** 2
  n_eff = B * 150 / 4.0
  assert np.all(np.abs(s.mean(0) - mean) < 5 * np.sqrt(var / n_eff))
  np.testing.assert_allclose(s.var(0), var, rtol=0.05)


QUADRATIC_SRC = r'''
__device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data) {
  const float y = x[0] - 0.2f;
  g[0] = -2.0f * y;
  return -y * y;
}
'''


def test_hmc_run_held_by_the_reference_notebook(tfp):
  """tests/golden/tf_notebook_hmc.json (TFP_Release_Notebook_0_11_0.ipynb, executed cell 36): the reference's own
  sample_chain(5, zeros([3]), HamiltonianMonteCarlo(lambda x: -(x - .2)**2, step_size=1., num_leapfrog_steps=2),
  num_burnin_steps=100, seed=(1, 2)) run.  Two unit leapfrogs map x -> 0.4 - x, m -> -m for any momentum, so the
  trajectory does not depend on the generator: the CUDA transitions (fused sample_chain over a run-time-compiled
  target) reproduce the reference's printed states, log-probs and (vanishing) acceptance corrections."""
  import json
  g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'tf_notebook_hmc.json')))
  tg = tfp.targets.UserTarget(1, QUADRATIC_SRC)
  k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=g['step_size'], num_leapfrog_steps=g['num_leapfrog_steps'])
  r = tfp.mcmc.sample_chain(g['num_results'], torch.zeros(3, 1, device=dev()), kernel=k,
                            num_burnin_steps=g['num_burnin_steps'], trace_fn=lambda _, kr: kr, seed=(1, 2))
  np.testing.assert_allclose(r.all_states[:, :, 0].cpu().numpy(), g['all_states'], atol=1e-5)
  assert bool(r.trace.is_accepted.all())
  np.testing.assert_allclose(r.trace.accepted_results.target_log_prob[0].cpu().numpy(),
                             g['target_log_prob_first_result'], atol=2e-6)
  assert float(r.trace.accepted_results.log_acceptance_correction.abs().max()) < 2e-6
