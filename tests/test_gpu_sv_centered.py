"""GPU parity of the CENTRED stochastic-volatility target (inference_gym/targets/stochastic_volatility.py:39-111;
SURVEY 8a row T4, second variant) against the oracle: log-prob / gradient (both coordinate systems, T = 60 and the
benchmark's T = 2516), a leapfrog trajectory within 1e-5 relative, NUTS trees, the TransformedTransitionKernel over the
constrained form == the kernel on the folded form, and the change-of-variables identity with the non-centred model."""
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


def t(a):
  return torch.tensor(np.asarray(a), device=dev())


def rel_err(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return np.max(np.abs(a - b) / (np.abs(b) + 1e-3 * np.max(np.abs(b)) + 1e-30))


def _state(T, B, seed, folded=True):
  rng = np.random.default_rng(seed)
  u = (0.3 * rng.standard_normal((B, T + 3))).astype(np.float32)
  u[:, 1] += 2.0
  u[:, 3:] += 2.0            # log-volatilities around the mean
  if not folded:
    u[:, 0] = np.tanh(u[:, 0])
    u[:, 2] = np.abs(u[:, 2]) + 0.2
  return u


@pytest.mark.parametrize('T,B', [(60, 33), (2516, 9)])
@pytest.mark.parametrize('folded', [True, False])
def test_centered_sv_logp_grad_matches_oracle(tfp, T, B, folded):
  y = tfp.targets.synthetic_sv_returns(T, seed=2)
  tg = (tfp.targets.StochasticVolatilityCentered if folded else tfp.targets.StochasticVolatilityCenteredConstrained)(y)
  o32 = otargets.StochasticVolatilityCentered(y, folded=folded)
  o64 = otargets.StochasticVolatilityCentered(y.astype(np.float64), dtype=np.float64, folded=folded)
  u = _state(T, B, 1, folded)
  lp, g = tg.log_prob_and_grad(t(u))
  lp64, g64 = o64.logp_grad(u.astype(np.float64))
  lp32, g32 = o32.logp_grad(u)
  assert rel_err(lp.cpu().numpy(), lp64) < max(5 * rel_err(lp32, lp64), 2e-6)
  scale = np.abs(g64).max(1, keepdims=True)
  err = np.max(np.abs(g.cpu().numpy() - g64) / scale)
  base = np.max(np.abs(g32 - g64) / scale)
  print('centred SV T=%d folded=%s: grad rel err %.2e (float32 oracle %.2e)' % (T, folded, err, base))
  assert err < max(5 * base, 5e-6)


def test_centered_sv_leapfrog_trajectory_1e5(tfp):
  from probability_b200 import _lib
  T, B, eps, L = 60, 33, 0.01, 4
  y = tfp.targets.synthetic_sv_returns(T, seed=2)
  tg = tfp.targets.StochasticVolatilityCentered(y)
  o32 = otargets.StochasticVolatilityCentered(y)
  o64 = otargets.StochasticVolatilityCentered(y.astype(np.float64), dtype=np.float64)
  x = _state(T, B, 3)
  m = np.random.default_rng(7).standard_normal(x.shape).astype(np.float32)
  lp0, g0 = o32.logp_grad(x)
  ctx = _lib.Context.get(dev()); ctx.bind_stream()
  xm, xx, xl, xg = t(m), t(x), t(lp0), t(g0)
  step = torch.tensor([eps], device=dev())
  om, ox, og, ol = torch.empty_like(xm), torch.empty_like(xx), torch.empty_like(xg), torch.empty_like(xl)
  _lib.check(ctx.lib.pb2_leapfrog(ctx.handle, tg.handle(ctx), B, _lib.ptr(xm), _lib.ptr(xx), _lib.ptr(xl), _lib.ptr(xg),
                                  _lib.ptr(step), 0, L, _lib.ptr(om), _lib.ptr(ox), _lib.ptr(ol), _lib.ptr(og)), ctx.handle)
  e = np.float64(eps)
  lp64, g64 = o64.logp_grad(x.astype(np.float64))
  v = m.astype(np.float64) + 0.5 * e * g64
  xx64, gg = x.astype(np.float64), g64
  for _ in range(L):
    xx64 = xx64 + e * v
    ll, gg = o64.logp_grad(xx64)
    v = v + e * gg
  m64 = v - 0.5 * e * gg
  r32 = omcmc.leapfrog(o32, m, x, lp0, g0, np.full(x.shape, eps, np.float32), L)
  assert rel_err(ox.cpu().numpy(), xx64) < max(1e-5, 3 * rel_err(r32[1], xx64))
  assert rel_err(om.cpu().numpy(), m64) < max(1e-5, 3 * rel_err(r32[0], m64))
  assert rel_err(ol.cpu().numpy(), ll) < 1e-5


def test_centered_sv_nuts_matches_oracle_and_transformed_kernel(tfp):
  T, B = 60, 48
  y = tfp.targets.synthetic_sv_returns(T, seed=2)
  folded = tfp.targets.StochasticVolatilityCentered(y)
  cons = tfp.targets.StochasticVolatilityCenteredConstrained(y)
  o32 = otargets.StochasticVolatilityCentered(y)
  u = _state(T, B, 5)
  parts = lambda a: [t(a[:, 0]), t(a[:, 1]), t(a[:, 2]), t(a[:, 3:])]
  k = tfp.mcmc.NoUTurnSampler(folded, step_size=0.02, max_tree_depth=5)
  seed = orng.key(41)
  st, kr = k.one_step(parts(u), k.bootstrap_results(parts(u)), seed=seed)
  lp0, g0 = o32.logp_grad(u)
  ref = omcmc.nuts_one_step(o32, u, lp0, g0, 0.02, seed, max_tree_depth=5)
  nl = kr.leapfrogs_taken.cpu().numpy()
  same = nl == ref['leapfrogs_taken']
  print('centred SV NUTS: identical trees %.4f (mean leapfrogs %.1f)' % (same.mean(), nl.mean()))
  assert same.mean() >= 0.97
  got = torch.cat([s.reshape(B, -1) for s in st], 1).cpu().numpy()
  np.testing.assert_allclose(got[same], ref['state'][same], rtol=2e-3, atol=2e-3)
  # the constrained form behind its bijectors takes the same transition
  bij = cons.default_event_space_bijector()
  ttk = tfp.mcmc.TransformedTransitionKernel(tfp.mcmc.NoUTurnSampler(cons, step_size=0.02, max_tree_depth=5), bij)
  x0 = [b.forward(p) for b, p in zip(bij, parts(u))]
  xs, kr_t = ttk.one_step(x0, ttk.bootstrap_results(x0), seed=seed)
  np.testing.assert_array_equal(kr_t.inner_results.leapfrogs_taken.cpu().numpy(), nl)
  for a, b_ in zip(kr_t.transformed_state, st):
    np.testing.assert_allclose(a.cpu().numpy(), b_.cpu().numpy(), rtol=1e-4, atol=1e-4)
  assert (xs[0].abs() < 1).all() and (xs[2] > 0).all()


def test_centered_and_non_centered_targets_agree_through_the_change_of_variables(tfp):
  """log p_nc(u_phi, m, u_s, z) = log p_c(u_phi, m, u_s, x(z)) + T log s - 1/2 log(1 - phi^2), both evaluated on the
  GPU: ties the centred kernel to the non-centred one (whose posterior is pinned on the Stan ground truth)."""
  T, B = 200, 17
  y = tfp.targets.synthetic_sv_returns(T, seed=4)
  nc = tfp.targets.StochasticVolatility(y)
  ce = tfp.targets.StochasticVolatilityCentered(y)
  u = (0.4 * np.random.default_rng(3).standard_normal((B, T + 3))).astype(np.float32)
  onc = otargets.StochasticVolatility(y.astype(np.float64), dtype=np.float64)
  phi, m, s, z = onc.constrain(u.astype(np.float64))
  h = np.empty_like(z)
  h[:, 0] = s * z[:, 0] / np.sqrt(1 - phi ** 2)
  for i in range(1, T):
    h[:, i] = phi * h[:, i - 1] + s * z[:, i]
  v = u.copy()
  v[:, 3:] = (h + m[:, None]).astype(np.float32)
  logdet = T * np.log(s) - 0.5 * np.log(1 - phi ** 2)
  lp_nc = nc.log_prob_and_grad(t(u))[0].cpu().numpy()
  lp_c = ce.log_prob_and_grad(t(v))[0].cpu().numpy()
  np.testing.assert_allclose(lp_nc, lp_c + logdet, rtol=2e-5, atol=2e-3)
