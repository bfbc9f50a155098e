"""Oracle pins of the CENTRED stochastic-volatility model (inference_gym/targets/stochastic_volatility.py:39-111):
its analytic gradient against torch.distributions + autograd in float64 (both coordinate systems), and its log-density
against the non-centred model's through the change of variables x = m + h(z) (same posterior, test infrastructure only)."""
import numpy as np
import pytest

from oracle import targets as otargets

torch = pytest.importorskip('torch')


@pytest.mark.parametrize('folded', [True, False])
def test_centered_sv_gradient_vs_autograd(folded):
  T = 40
  y = otargets.synthetic_sv_returns(T, seed=1).astype(np.float64)
  tg = otargets.StochasticVolatilityCentered(y, dtype=np.float64, folded=folded)
  rng = np.random.default_rng(0)
  u = 0.3 * rng.standard_normal((4, T + 3))
  u[:, 1] += 2.0
  if not folded:
    u[:, 0] = np.tanh(u[:, 0])
    u[:, 2] = np.abs(u[:, 2]) + 0.2
  lp, g = tg.logp_grad(u)
  D = torch.distributions
  f64 = lambda v: torch.tensor(v, dtype=torch.float64)
  ut = torch.tensor(u, requires_grad=True)
  sp = torch.nn.functional.softplus
  phi = 2 * torch.sigmoid(ut[:, 0]) - 1 if folded else ut[:, 0]
  s = sp(ut[:, 2]) if folded else ut[:, 2]
  m, x = ut[:, 1], ut[:, 3:]
  l = (D.Beta(f64(20.), f64(1.5)).log_prob((phi + 1) / 2) - np.log(2.) + D.Cauchy(f64(0.), f64(5.)).log_prob(m)
       + D.HalfCauchy(f64(2.)).log_prob(s) + D.Normal(m, s / torch.sqrt(1 - phi ** 2)).log_prob(x[:, 0])
       + D.Normal(m[:, None] + phi[:, None] * (x[:, :-1] - m[:, None]), s[:, None]).log_prob(x[:, 1:]).sum(1)
       + D.Normal(f64(0.), torch.exp(x / 2)).log_prob(f64(y)[None, :]).sum(1))
  if folded:
    l = l + (np.log(2.) - sp(-ut[:, 0]) - sp(ut[:, 0])) + (-sp(-ut[:, 2]))
  l.sum().backward()
  np.testing.assert_allclose(lp, l.detach().numpy(), rtol=1e-10)
  np.testing.assert_allclose(g, ut.grad.numpy(), rtol=1e-9, atol=1e-11)


def test_centered_and_non_centered_densities_agree_through_the_change_of_variables():
  """x_t = m + h_t(z): log p_nc(phi, m, s, z) = log p_c(phi, m, s, x(z)) + log |dx/dz|, dx/dz lower triangular with
  diagonal (s / q, s, .., s)."""
  T = 30
  y = otargets.synthetic_sv_returns(T, seed=2).astype(np.float64)
  nc = otargets.StochasticVolatility(y, dtype=np.float64)
  ce = otargets.StochasticVolatilityCentered(y, dtype=np.float64, folded=True)
  u = 0.4 * np.random.default_rng(3).standard_normal((5, T + 3))
  phi, m, s, z = nc.constrain(u)
  h = np.empty_like(z)
  h[:, 0] = s * z[:, 0] / np.sqrt(1 - phi ** 2)
  for t in range(1, T):
    h[:, t] = phi * h[:, t - 1] + s * z[:, t]
  v = u.copy()
  v[:, 3:] = h + m[:, None]
  logdet = T * np.log(s) - 0.5 * np.log(1 - phi ** 2)
  np.testing.assert_allclose(nc.logp_grad(u)[0], ce.logp_grad(v)[0] + logdet, rtol=1e-10)
