"""GPU tests of TransformedTransitionKernel and the event-space bijectors (SURVEY 8f-1, row W1;
tfp/mcmc/transformed_kernel.py:86-140,165-440; bijectors/{identity,exp,softplus,sigmoid}.py)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu

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


def _np_bijector(kind, lo, hi, u):
  """(forward, d forward / du, fldj, d fldj / du) in float64."""
  sp = lambda v: np.logaddexp(0.0, v)
  sg = lambda v: 1.0 / (1.0 + np.exp(-v))
  if kind == 0:
    return u, np.ones_like(u), np.zeros_like(u), np.zeros_like(u)
  if kind == 1:
    return np.exp(u), np.exp(u), u, np.ones_like(u)
  if kind == 2:
    return sp(u), sg(u), -sp(-u), sg(-u)
  w = hi - lo
  return lo + w * sg(u), w * sg(u) * sg(-u), np.log(w) - sp(-u) - sp(u), sg(-u) - sg(u)


def test_bijector_host_maps(tfp):
  b = tfp.bijectors
  x = torch.linspace(-4, 4, 41, device=dev(), dtype=torch.float64)
  for bij in (b.Identity(), b.Exp(), b.Softplus(), b.Sigmoid(-1., 1.), b.Sigmoid(2., 7.5)):
    xx = x.clone().requires_grad_(True)
    y = bij.forward(xx)
    np.testing.assert_allclose(bij.inverse(y).detach().cpu().numpy(), x.cpu().numpy(), rtol=1e-9, atol=1e-9)
    (dy,) = torch.autograd.grad(y.sum(), xx)
    np.testing.assert_allclose(bij.forward_log_det_jacobian(x).cpu().numpy(), dy.log().cpu().numpy(), rtol=1e-9, atol=1e-9)
  assert tuple(b.Exp().forward_log_det_jacobian(torch.ones(3, 4), event_ndims=1).shape) == (3,)


def test_transformed_target_logp_grad_all_bijectors(tfp):
  """lp(u) = lp_x(b(u)) + sum fldj(u), grad_u = grad_x b'(u) + fldj'(u) for every bijector kind (float64 reference)."""
  b = tfp.bijectors
  tg = tfp.targets.EightSchools()          # parts [1, 1, 8]: any fused target will do for the arithmetic
  o64 = otargets.EightSchools(dtype=np.float64)
  bijs = [b.Sigmoid(-3., 5.), b.Softplus(), b.Exp()]
  tt = tfp.mcmc.transformed_kernel.TransformedTarget(tg, bijs)
  rng = np.random.default_rng(0)
  u = (0.6 * rng.standard_normal((200, 10))).astype(np.float32)
  lp, g = tt.log_prob_and_grad(t(u))
  kinds = [3] + [2] + [1] * 8
  x = np.empty((200, 10)); db = np.empty_like(x); lj = np.empty_like(x); dj = np.empty_like(x)
  for d, k in enumerate(kinds):
    x[:, d], db[:, d], lj[:, d], dj[:, d] = _np_bijector(k, -3.0, 5.0, u[:, d].astype(np.float64))
  lpx, gx = o64.logp_grad(x)
  np.testing.assert_allclose(lp.cpu().numpy(), lpx + lj.sum(1), rtol=2e-5, atol=2e-4)
  ref_g = gx * db + dj
  scale = np.abs(ref_g).max(1, keepdims=True)
  assert np.max(np.abs(g.cpu().numpy() - ref_g) / scale) < 2e-5
  with pytest.raises(ValueError):
    tfp.mcmc.transformed_kernel.TransformedTarget(tg, bijs[:2])
  with pytest.raises(TypeError):
    tfp.mcmc.TransformedTransitionKernel(tfp.mcmc.NoUTurnSampler(tg, 0.1), bijector=lambda x: x)


def _sv(tfp, T=60):
  yv = tfp.targets.synthetic_sv_returns(T=T, seed=2)
  rng = np.random.default_rng(4)
  u = (0.3 * rng.standard_normal((48, T + 3))).astype(np.float32)
  u[:, 0] += 2.0
  u[:, 1] += 5.0
  return yv, u


def _parts(a):
  a = t(a)
  return [a[:, 0].contiguous(), a[:, 1].contiguous(), a[:, 2].contiguous(), a[:, 3:].contiguous()]


def test_constrained_model_through_bijectors_equals_folded_target(tfp):
  """The stochastic-volatility model in its own coordinates + its default event-space bijector == the target with the
  bijector folded in (what round 1 shipped) == the oracle: log-prob and gradient."""
  yv, u = _sv(tfp)
  folded = tfp.targets.StochasticVolatility(yv)
  cons = tfp.targets.StochasticVolatilityConstrained(yv)
  tt = tfp.mcmc.transformed_kernel.TransformedTarget(cons, cons.default_event_space_bijector())
  lp_a, g_a = tt.log_prob_and_grad(t(u))
  lp_b, g_b = folded.log_prob_and_grad(t(u))
  np.testing.assert_allclose(lp_a.cpu().numpy(), lp_b.cpu().numpy(), rtol=2e-6, atol=1e-4)
  np.testing.assert_allclose(g_a.cpu().numpy(), g_b.cpu().numpy(), rtol=1e-4, atol=1e-4)
  lp64, g64 = otargets.StochasticVolatility(yv.astype(np.float64), dtype=np.float64).logp_grad(u.astype(np.float64))
  np.testing.assert_allclose(lp_a.cpu().numpy(), lp64, rtol=2e-6, atol=1e-4)
  # the constrained log-prob itself: at x = forward(u) it is lp(u) - fldj(u)
  x = u.astype(np.float64).copy()
  x[:, 0], _, lj0, _ = _np_bijector(3, -1.0, 1.0, u[:, 0].astype(np.float64))
  x[:, 2], _, lj2, _ = _np_bijector(2, 0, 1, u[:, 2].astype(np.float64))
  lp_c, _ = cons.log_prob_and_grad(t(x.astype(np.float32)))
  np.testing.assert_allclose(lp_c.cpu().numpy(), lp64 - lj0 - lj2, rtol=5e-6, atol=5e-4)


@pytest.mark.parametrize('sampler', ['nuts', 'hmc'])
def test_transformed_kernel_equals_kernel_on_folded_target(tfp, sampler):
  yv, u = _sv(tfp)
  folded = tfp.targets.StochasticVolatility(yv)
  cons = tfp.targets.StochasticVolatilityConstrained(yv)
  bij = cons.default_event_space_bijector()
  mk = (lambda tg: tfp.mcmc.NoUTurnSampler(tg, step_size=0.03, max_tree_depth=5)) if sampler == 'nuts' else \
      (lambda tg: tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.02, num_leapfrog_steps=4))
  ttk = tfp.mcmc.TransformedTransitionKernel(mk(cons), bij)
  plain = mk(folded)
  x0 = [b.forward(p) for b, p in zip(bij, _parts(u))]            # constrained initial state
  kr_t = ttk.bootstrap_results(x0)
  for a, b_ in zip(kr_t.transformed_state, _parts(u)):
    np.testing.assert_allclose(a.cpu().numpy(), b_.cpu().numpy(), rtol=2e-5, atol=2e-5)
  kr_t = ttk.bootstrap_results(transformed_init_state=_parts(u))
  kr_p = plain.bootstrap_results(_parts(u))
  seed = orng.key(41)
  xs, kr_t1 = ttk.one_step(x0, kr_t, seed=seed)
  us, kr_p1 = plain.one_step(_parts(u), kr_p, seed=seed)
  if sampler == 'nuts':
    np.testing.assert_array_equal(kr_t1.inner_results.leapfrogs_taken.cpu().numpy(), kr_p1.leapfrogs_taken.cpu().numpy())
  else:
    np.testing.assert_array_equal(kr_t1.inner_results.is_accepted.cpu().numpy(), kr_p1.is_accepted.cpu().numpy())
  for a, b_ in zip(kr_t1.transformed_state, us):
    np.testing.assert_allclose(a.cpu().numpy(), b_.cpu().numpy(), rtol=1e-4, atol=1e-4)
  for bj, xc, uc in zip(bij, xs, us):                            # reported states are in the constrained space
    np.testing.assert_allclose(xc.cpu().numpy(), bj.forward(uc).cpu().numpy(), rtol=1e-4, atol=1e-4)
  assert (xs[0].abs() < 1).all() and (xs[2] > 0).all()
  with pytest.raises(ValueError):
    ttk.bootstrap_results()
  # sample_chain: the fused run (one pb2_run in the unconstrained space) equals the step loop
  fld = (lambda kr: kr.inner_results.leapfrogs_taken) if sampler == 'nuts' else (lambda kr: kr.inner_results.is_accepted)
  a = tfp.mcmc.sample_chain(5, x0, previous_kernel_results=kr_t, kernel=ttk, seed=8, trace_fn=lambda _, kr: fld(kr))
  b2 = tfp.mcmc.sample_chain(5, x0, previous_kernel_results=kr_t, kernel=ttk, seed=8,
                             trace_fn=lambda _, kr: fld(kr).int() + 0)     # computes on values: step loop
  np.testing.assert_array_equal(a.trace.int().cpu().numpy(), b2.trace.cpu().numpy())
  for p, q in zip(a.all_states, b2.all_states):
    np.testing.assert_allclose(p.cpu().numpy(), q.cpu().numpy(), rtol=1e-5, atol=1e-6)
  assert (a.all_states[0].abs() < 1).all() and (a.all_states[2] > 0).all()
