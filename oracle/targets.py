"""Oracle targets: closed-form log-densities + analytic gradients (TEST INFRASTRUCTURE).

Each class exposes `dim`, `logp_grad(x[B,D]) -> (lp[B], g[B,D])` in the requested
dtype (float32 mirrors the reference arithmetic; float64 is the shadow used to
validate the gradients against torch.autograd in tests).
"""
import numpy as np

_HALF_LOG_2PI = 0.9189385332046727


def _softplus(x):
  # tfp/internal/backend/numpy/numpy_math.py:981-987
  return np.log1p(np.exp(-np.abs(x))) + np.maximum(x, 0)


def _sigmoid(x):
  return np.where(x >= 0, 1.0 / (1.0 + np.exp(-np.abs(x))),
                  np.exp(-np.abs(x)) / (1.0 + np.exp(-np.abs(x)))).astype(x.dtype)


class EightSchools:
  """Non-centred Eight Schools. tfp/mcmc/eight_schools_hmc.py:41-57 (model),
  :69-76 (data); Normal.log_prob formula distributions/normal.py:182-188.
  State layout: x = [mu, tau, z_0..z_{J-1}]  (parts [], [], [J])."""
  Y = np.array([28, 8, -3, 7, -1, 1, 18, 12], np.float64)
  SIGMA = np.array([15, 10, 16, 11, 9, 11, 10, 18], np.float64)

  def __init__(self, y=None, sigma=None, dtype=np.float32):
    self.dtype = dtype
    self.y = np.asarray(self.Y if y is None else y, dtype)
    self.sigma = np.asarray(self.SIGMA if sigma is None else sigma, dtype)
    self.J = self.y.size
    self.dim = self.J + 2
    self.part_sizes = [1, 1, self.J]

  def logp_grad(self, x):
    dt = self.dtype
    x = np.asarray(x, dt)
    mu, tau, z = x[:, 0], x[:, 1], x[:, 2:]
    c = dt(_HALF_LOG_2PI)
    e = np.exp(tau)
    loc = mu[:, None] + e[:, None] * z
    r = (self.y[None, :] - loc) / self.sigma[None, :]
    lp = (dt(-0.5) * (mu / dt(10)) ** 2 - (c + np.log(dt(10)))
          + dt(-0.5) * (tau - dt(5)) ** 2 - c
          + np.sum(dt(-0.5) * z * z - c, axis=1)
          + np.sum(dt(-0.5) * r * r - (c + np.log(self.sigma))[None, :], axis=1))
    w = r / self.sigma[None, :]
    g = np.empty_like(x)
    g[:, 0] = -mu / dt(100) + np.sum(w, axis=1)
    g[:, 1] = -(tau - dt(5)) + e * np.sum(w * z, axis=1)
    g[:, 2:] = -z + e[:, None] * w
    return lp.astype(dt), g.astype(dt)


def ill_conditioned_covariance(ndims=100, gamma_shape_parameter=0.5, seed=10):
  """inference_gym/targets/ill_conditioned_gaussian.py:66-75 (NumPy only)."""
  rng = np.random.RandomState(seed=seed & (2**32 - 1))
  eigenvalues = 1. / np.sort(rng.gamma(shape=gamma_shape_parameter, scale=1., size=ndims))
  q, r = np.linalg.qr(rng.randn(ndims, ndims))
  q *= np.sign(np.diag(r))
  return (q * eigenvalues).dot(q.T), eigenvalues


def gaussian_precision_from_cov(cov):
  """Fix L32 = chol(float32(cov)); P = (L32 L32^T)^-1 in float64, rounded to fp32;
  log-normaliser const = -sum(log diag L32) - D/2 log 2pi  (the
  precision matrix shared by oracle and kernel)."""
  cov32 = np.asarray(cov, np.float32)
  L = np.linalg.cholesky(cov32.astype(np.float64)).astype(np.float32).astype(np.float64)
  P = np.linalg.inv(L @ L.T)
  P = 0.5 * (P + P.T)
  const = -np.sum(np.log(np.diag(L))) - 0.5 * cov32.shape[0] * np.log(2 * np.pi)
  return P.astype(np.float32), np.float32(const)


class DenseGaussian:
  """MVNTriL(loc, chol(cov)).log_prob  (ill_conditioned_gaussian.py:77-81,105-106;
  mvn_linear_operator.py:242) evaluated as lp = -1/2 (x-mu)^T P (x-mu) + const, g = -P(x-mu)."""

  def __init__(self, precision, const=0.0, loc=None, dtype=np.float32):
    self.dtype = dtype
    self.P = np.asarray(precision, dtype)
    self.dim = self.P.shape[0]
    self.const = dtype(const)
    self.loc = np.zeros(self.dim, dtype) if loc is None else np.asarray(loc, dtype)
    self.part_sizes = [self.dim]

  def logp_grad(self, x):
    dt = self.dtype
    xc = np.asarray(x, dt) - self.loc[None, :]
    g = -(xc @ self.P.T).astype(dt)
    lp = dt(0.5) * np.sum(xc * g, axis=1) + self.const
    return lp.astype(dt), g.astype(dt)


class LogisticRegression:
  """inference_gym/targets/logistic_regression.py:88-103 (+bias :36-39);
  prior+likelihood bayesian_model.py:100-102; Bernoulli-logits bernoulli.py:119-135.
  `features` here ALREADY include the trailing ones column."""

  def __init__(self, features, labels, dtype=np.float32):
    self.dtype = dtype
    self.X = np.asarray(features, dtype)
    self.y = np.asarray(labels, dtype)
    self.dim = self.X.shape[1]
    self.part_sizes = [self.dim]

  def logp_grad(self, x):
    dt = self.dtype
    th = np.asarray(x, dt)
    z = (th @ self.X.T).astype(dt)                       # [B,N]
    ll = self.y[None, :] * z - _softplus(z)
    lp = (np.sum(dt(-0.5) * th * th - dt(_HALF_LOG_2PI), axis=1) + np.sum(ll, axis=1))
    w = self.y[None, :] - _sigmoid(z)
    g = -th + (w @ self.X).astype(dt)
    return lp.astype(dt), g.astype(dt)


class StochasticVolatility:
  """Non-centred vectorised stochastic volatility in UNCONSTRAINED space.
  inference_gym/targets/vectorized_stochastic_volatility.py:233-309 (model),
  :47-99 (the FFT convolution == the AR(1) recurrence below), :346-356 (bijectors);
  Beta beta.py:340-345, Cauchy cauchy.py:174-179, HalfCauchy half_cauchy.py:138-145,
  Sigmoid(lo,hi) sigmoid.py:145-179, Softplus softplus.py:139-172;
  TransformedTransitionKernel log-prob transform transformed_kernel.py:86-140.
  State layout: u = [u_phi, m, u_s, z_0..z_{T-1}]."""

  def __init__(self, centered_returns, dtype=np.float32):
    self.dtype = dtype
    self.y = np.asarray(centered_returns, dtype)
    self.T = self.y.size
    self.dim = self.T + 3
    self.part_sizes = [1, 1, 1, self.T]

  def constrain(self, u):
    u = np.asarray(u, self.dtype)
    phi = 2.0 * _sigmoid(u[:, 0]) - 1.0
    m = u[:, 1]
    s = _softplus(u[:, 2])
    return phi.astype(self.dtype), m, s.astype(self.dtype), u[:, 3:]

  def logp_grad(self, u):
    dt = self.dtype
    u = np.asarray(u, dt)
    B, T = u.shape[0], self.T
    u1, m, u3, z = u[:, 0], u[:, 1], u[:, 2], u[:, 3:]
    sg = _sigmoid(u1)
    sgm = _sigmoid(-u1)
    phi = dt(2) * sg - dt(1)
    s = _softplus(u3)
    omp2 = dt(1) - phi * phi
    rs = dt(1) / np.sqrt(omp2)
    # forward recurrence
    h = np.empty((B, T), dt)
    h[:, 0] = s * z[:, 0] * rs
    for t in range(1, T):
      h[:, t] = phi * h[:, t - 1] + s * z[:, t]
    hm = h + m[:, None]
    y2e = (self.y[None, :] ** 2) * np.exp(-hm)
    lik = np.sum(dt(-0.5) * y2e - dt(_HALF_LOG_2PI) - dt(0.5) * hm, axis=1)
    b = (phi + dt(1)) * dt(0.5)
    # Beta(20, 1.5).log_prob(b) - log 2 (Scale(2) Jacobian)
    from math import lgamma
    lbeta = lgamma(20.0) + lgamma(1.5) - lgamma(21.5)
    lp_phi = dt(19.0) * np.log(b) + dt(0.5) * np.log1p(-b) - dt(lbeta) - dt(np.log(2.0))
    lp_m = -dt(np.log(np.pi * 5.0)) - np.log1p((m / dt(5)) ** 2)
    lp_s = dt(np.log(2.0)) - dt(np.log(np.pi * 2.0)) - np.log1p((s / dt(2)) ** 2)
    lp_z = np.sum(dt(-0.5) * z * z - dt(_HALF_LOG_2PI), axis=1)
    # forward-log-det-Jacobians: Sigmoid(-1,1): log(2) - softplus(-u) - softplus(u); Softplus: -softplus(-u)
    fldj = (dt(np.log(2.0)) - _softplus(-u1) - _softplus(u1)) + (-_softplus(-u3))
    lp = lik + lp_phi + lp_m + lp_s + lp_z + fldj
    # adjoint (reverse recurrence)
    a = dt(0.5) * (y2e - dt(1))
    lam = np.empty((B, T), dt)
    lam[:, T - 1] = a[:, T - 1]
    for t in range(T - 2, -1, -1):
      lam[:, t] = a[:, t] + phi * lam[:, t + 1]
    g = np.empty_like(u)
    g[:, 3:] = s[:, None] * lam - z
    g[:, 3] = s * lam[:, 0] * rs - z[:, 0]
    d_m = np.sum(a, axis=1) - (dt(2) * m / dt(25)) / (dt(1) + (m / dt(5)) ** 2)
    c = z.copy()
    c[:, 0] = z[:, 0] * rs
    d_s = np.sum(lam * c, axis=1) - (s / dt(2)) / (dt(1) + (s / dt(2)) ** 2)
    d_phi = (np.sum(lam[:, 1:] * h[:, :-1], axis=1)
             + lam[:, 0] * s * z[:, 0] * phi * rs * rs * rs
             + dt(0.5) * (dt(19.0) / b - dt(0.5) / (dt(1) - b)))
    g[:, 0] = d_phi * (dt(2) * sg * sgm) + (sgm - sg)
    g[:, 1] = d_m
    g[:, 2] = d_s * _sigmoid(u3) + _sigmoid(-u3)
    return lp.astype(dt), g.astype(dt)


class StochasticVolatilityCentered:
  """CENTRED stochastic volatility: one latent log-volatility x_t per time step,
  inference_gym/targets/stochastic_volatility.py:39-52 (AR(1) prior x_0 ~ N(m, s / sqrt(1 - phi^2)),
  x_t ~ N(m + phi (x_{t-1} - m), s)), :68-90 (priors: 2 Beta(20, 1.5) - 1, Cauchy(0, 5), HalfCauchy(0, 2)),
  :102-111 (y_t ~ N(0, exp(x_t / 2))).  Same posterior as the non-centred model, its gradient is a 3-point stencil.
  State layout [phi, m, s, x_0..x_{T-1}]; `folded=True`: unconstrained coordinates u = [u_phi, m, u_s, x] with the
  Sigmoid(-1, 1) / Softplus bijectors and their forward-log-det-Jacobians folded in (transformed_kernel.py:86-140)."""

  def __init__(self, centered_returns, dtype=np.float32, folded=True):
    self.dtype = dtype
    self.y = np.asarray(centered_returns, dtype)
    self.T = self.y.size
    self.dim = self.T + 3
    self.part_sizes = [1, 1, 1, self.T]
    self.folded = folded

  def logp_grad(self, u):
    dt = self.dtype
    u = np.asarray(u, dt)
    u1, m, u3, x = u[:, 0], u[:, 1], u[:, 2], u[:, 3:]
    if self.folded:
      sg, sgm = _sigmoid(u1), _sigmoid(-u1)
      phi = dt(2) * sg - dt(1)
      s = _softplus(u3)
    else:
      phi, s = u1, u3
    q = np.sqrt(dt(1) - phi * phi)
    d = x - m[:, None]
    e = np.empty_like(x)
    e[:, 0] = d[:, 0] * q / s
    e[:, 1:] = (d[:, 1:] - phi[:, None] * d[:, :-1]) / s[:, None]
    y2e = (self.y[None, :] ** 2) * np.exp(-x)
    lp_x = np.sum(dt(-0.5) * e * e - dt(_HALF_LOG_2PI) - np.log(s)[:, None], axis=1) + np.log(q)
    lik = np.sum(dt(-0.5) * y2e - dt(_HALF_LOG_2PI) - dt(0.5) * x, axis=1)
    b = (phi + dt(1)) * dt(0.5)
    from math import lgamma
    lbeta = lgamma(20.0) + lgamma(1.5) - lgamma(21.5)
    lp_phi = dt(19.0) * np.log(b) + dt(0.5) * np.log1p(-b) - dt(lbeta) - dt(np.log(2.0))
    lp_m = -dt(np.log(np.pi * 5.0)) - np.log1p((m / dt(5)) ** 2)
    lp_s = dt(np.log(2.0)) - dt(np.log(np.pi * 2.0)) - np.log1p((s / dt(2)) ** 2)
    lp = lik + lp_x + lp_phi + lp_m + lp_s
    g = np.empty_like(u)
    c = np.ones_like(x)
    c[:, 0] = q
    gx = -e * c / s[:, None] + dt(0.5) * (y2e - dt(1))
    gx[:, :-1] += e[:, 1:] * (phi / s)[:, None]
    g[:, 3:] = gx
    cm = np.full_like(x, dt(1)) * (dt(1) - phi)[:, None]
    cm[:, 0] = q
    d_m = np.sum(e * cm, axis=1) / s - (dt(2) * m / dt(25)) / (dt(1) + (m / dt(5)) ** 2)
    d_s = np.sum(e * e - dt(1), axis=1) / s - (s / dt(2)) / (dt(1) + (s / dt(2)) ** 2)
    d_phi = (np.sum(e[:, 1:] * d[:, :-1], axis=1) / s + e[:, 0] * d[:, 0] * phi / (s * q) - phi / (q * q)
             + dt(0.5) * (dt(19.0) / b - dt(0.5) / (dt(1) - b)))
    if self.folded:
      lp = lp + (dt(np.log(2.0)) - _softplus(-u1) - _softplus(u1)) + (-_softplus(-u3))
      g[:, 0] = d_phi * (dt(2) * sg * sgm) + (sgm - sg)
      g[:, 2] = d_s * _sigmoid(u3) + _sigmoid(-u3)
    else:
      g[:, 0] = d_phi
      g[:, 2] = d_s
    g[:, 1] = d_m
    return lp.astype(dt), g.astype(dt)


def synthetic_sv_returns(T=2516, phi=0.95, s=0.25, m=None, seed=0):
  """Synthetic S&P500-shape centred returns simulated from the model itself
  (SURVEY section 8d, C4)."""
  rng = np.random.default_rng(seed)
  m = 2.0 * np.log(15.0) if m is None else m   # S&P abs. daily moves ~ 15 points
  h = np.empty(T)
  h[0] = s * rng.standard_normal() / np.sqrt(1 - phi * phi)
  for t in range(1, T):
    h[t] = phi * h[t - 1] + s * rng.standard_normal()
  y = rng.standard_normal(T) * np.exp(0.5 * (h + m))
  return (y - y.mean()).astype(np.float32)


def synthetic_logistic_data(n=1000, d=24, seed=0):
  """standardised N(0,1) features + ones column, theta*~N(0,1)."""
  rng = np.random.default_rng(seed)
  X = rng.standard_normal((n, d))
  X = (X - X.mean(0)) / X.std(0)
  X = np.concatenate([X, np.ones((n, 1))], axis=1)
  theta = rng.standard_normal(d + 1)
  p = 1.0 / (1.0 + np.exp(-(X @ theta)))
  y = (rng.random(n) < p).astype(np.float32)
  return X.astype(np.float32), y
