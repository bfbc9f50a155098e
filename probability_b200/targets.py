"""Named target densities with fused log-prob + gradient CUDA kernels.

A target object plays the role of the reference's `target_log_prob_fn` callable
(hmc.py:413-415): calling it with the state parts returns the per-chain log-prob.
Because the transition kernels are persistent CUDA kernels, an arbitrary Python
callable cannot be used; the kernels accept these target objects only and raise
otherwise (no CPU / autodiff fallback).

  EightSchools            tfp/mcmc/eight_schools_hmc.py:41-76
  DenseGaussian           MVNTriL log_prob; IllConditionedGaussian =
                          inference_gym/targets/ill_conditioned_gaussian.py:30-110
  LogisticRegression      inference_gym/targets/logistic_regression.py:42-171
  StochasticVolatility    inference_gym/targets/vectorized_stochastic_volatility.py:102-440
                          (non-centred, unconstrained space; `constrain` maps back)
  StochasticVolatilityCentered   inference_gym/targets/stochastic_volatility.py:39-111 (one latent per time step)
  UserTarget              any density, given as CUDA source for one chain's log-prob + gradient (compiled at run time
                          into the same chain kernels; csrc/pb2_user_target.cuh)
"""
import ctypes as C

import numpy as np

from probability_b200 import _lib


class Target:
  """Base: owns a pb2_target handle per device, created lazily."""
  kind = None

  def __init__(self, dim, n_rows, a, b=None, scalar=0.0, part_sizes=None):
    self.dim = int(dim)
    self.n_rows = int(n_rows)
    self._a = np.ascontiguousarray(a, np.float32)
    self._b = None if b is None else np.ascontiguousarray(b, np.float32)
    self._scalar = float(scalar)
    self.part_sizes = list(part_sizes) if part_sizes is not None else [self.dim]
    self._handles = {}

  def handle(self, ctx):
    h = self._handles.get(ctx.device_index)
    if h is None:
      desc = _lib.TargetDesc(
          kind=self.kind, dim=self.dim, n_rows=self.n_rows,
          h_a=self._a.ctypes.data_as(_lib.c_f32p),
          h_b=(self._b.ctypes.data_as(_lib.c_f32p) if self._b is not None else None),
          scalar=self._scalar)
      h = C.c_void_p()
      _lib.check(ctx.lib.pb2_target_create(ctx.handle, C.byref(desc), C.byref(h)), ctx.handle)
      self._handles[ctx.device_index] = h
    return h

  # -- the `target_log_prob_fn` protocol ------------------------------------
  def log_prob_and_grad(self, x):
    """x: float32 CUDA tensor [B, D] -> (lp [B], grad [B, D])  (mcmc/internal/util.py:286-308)."""
    import torch
    if x.dim() != 2 or x.shape[1] != self.dim:
      raise ValueError('expected state of shape [chains, {}], got {}'.format(self.dim, tuple(x.shape)))
    x = x.contiguous().float()
    ctx = _lib.Context.get(x.device)
    ctx.bind_stream()
    lp = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    g = torch.empty_like(x)
    _lib.check(ctx.lib.pb2_logp_grad(ctx.handle, self.handle(ctx), x.shape[0], _lib.ptr(x), _lib.ptr(lp),
                                     _lib.ptr(g)), ctx.handle)
    return lp, g

  def __call__(self, *state_parts):
    from probability_b200.mcmc import _engine
    x, _, _ = _engine.flatten_state(list(state_parts))
    return self.log_prob_and_grad(x)[0]


class UserTarget(Target):
  """A user-defined target density: the engine's form of the reference's arbitrary `target_log_prob_fn`
  (tfp/mcmc/hmc.py:413-415; value and gradient as in mcmc/internal/util.py:246-308).

  `source` is CUDA C++ defining, for ONE chain,

      __device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data);

  which returns log p(x) up to a constant and writes the gradient into g[0 .. dim).  It is compiled with NVRTC for
  sm_100a into the leapfrog / HMC / NUTS chain kernels of the named targets (warp per chain), so every kernel of
  `tfp.mcmc` -- `sample_chain`, step-size adaptation, diagnostics -- runs on it unchanged.  `data` (optional) is a flat
  float32 array handed to every call; `part_sizes` splits the flat state into the state parts `sample_chain` sees.
  With `cooperative=True` the function takes a trailing `int lane` and is called by all 32 lanes of the chain's warp
  (every lane writes a disjoint part of g, all lanes return the same value; `pb2::warp_sum` is available).
  A compile error raises `Pb2Error` carrying the compiler's log (line numbers refer to `source`)."""
  kind = _lib.TARGET_USER

  def __init__(self, dim, source, data=None, part_sizes=None, cooperative=False):
    if not isinstance(source, str) or 'target_log_prob_and_grad' not in source:
      raise TypeError('`source` must be CUDA source that defines target_log_prob_and_grad (see UserTarget docs)')
    dim = int(dim)
    if not 1 <= dim <= 256:
      raise ValueError('UserTarget: 1 <= dim <= 256, got {}'.format(dim))
    data = np.zeros([0], np.float32) if data is None else np.ascontiguousarray(np.asarray(data, np.float32).reshape(-1))
    if part_sizes is not None and sum(int(n) for n in part_sizes) != dim:
      raise ValueError('part_sizes must sum to dim')
    self.source = source
    self.flags = 1 if cooperative else 0   # PB2_USER_COOPERATIVE
    super().__init__(dim=dim, n_rows=data.size, a=data, part_sizes=part_sizes)

  def check(self):
    """Compile only (no GPU needed); raises Pb2Error with the compiler's log on failure."""
    lib = _lib.load()
    rc = lib.pb2_user_target_check(self.source.encode(), self.dim, self.flags, _lib.CSRC_DIR.encode())
    _lib.check(rc)
    return True

  def handle(self, ctx):
    h = self._handles.get(ctx.device_index)
    if h is None:
      h = C.c_void_p()
      data = self._a.ctypes.data_as(_lib.c_f32p) if self._a.size else None
      _lib.check(ctx.lib.pb2_target_create_user(ctx.handle, self.dim, self.source.encode(), self.flags, data, self._a.size,
                                                _lib.CSRC_DIR.encode(), C.byref(h)), ctx.handle)
      self._handles[ctx.device_index] = h
    return h


class EightSchools(Target):
  kind = _lib.TARGET_EIGHT_SCHOOLS
  TREATMENT_EFFECTS = [28, 8, -3, 7, -1, 1, 18, 12]      # eight_schools_hmc.py:69-72
  TREATMENT_STDDEVS = [15, 10, 16, 11, 9, 11, 10, 18]    # :73-76

  def __init__(self, treatment_effects=None, treatment_stddevs=None):
    y = self.TREATMENT_EFFECTS if treatment_effects is None else treatment_effects
    s = self.TREATMENT_STDDEVS if treatment_stddevs is None else treatment_stddevs
    y = np.asarray(y, np.float32)
    s = np.asarray(s, np.float32)
    if y.shape != s.shape or y.ndim != 1:
      raise ValueError('treatment_effects and treatment_stddevs must be 1-d and equal length')
    super().__init__(dim=y.size + 2, n_rows=y.size, a=y, b=s, part_sizes=[1, 1, y.size])


class DenseGaussian(Target):
  """N(loc, covariance) evaluated as -1/2 (x-loc)^T P (x-loc) + const with a fixed fp32
  precision matrix P (cholesky of float32(cov), inverted in float64)."""
  kind = _lib.TARGET_DENSE_GAUSSIAN

  def __init__(self, covariance=None, loc=None, precision=None, log_normalizer=None):
    if precision is None:
      cov32 = np.asarray(covariance, np.float32)
      L = np.linalg.cholesky(cov32.astype(np.float64)).astype(np.float32).astype(np.float64)
      P = np.linalg.inv(L @ L.T)
      P = 0.5 * (P + P.T)
      const = -np.sum(np.log(np.diag(L))) - 0.5 * cov32.shape[0] * np.log(2 * np.pi)
    else:
      P = np.asarray(precision, np.float64)
      const = 0.0 if log_normalizer is None else log_normalizer
    self.precision = P.astype(np.float32)
    self.log_normalizer = float(const)
    d = P.shape[0]
    self.loc = np.zeros(d, np.float32) if loc is None else np.asarray(loc, np.float32)
    super().__init__(dim=d, n_rows=d, a=self.precision, b=self.loc, scalar=const)


class IllConditionedGaussian(DenseGaussian):
  def __init__(self, ndims=100, gamma_shape_parameter=0.5, max_eigvalue=None, seed=10):
    rng = np.random.RandomState(seed=seed & (2**32 - 1))
    eigenvalues = 1. / np.sort(rng.gamma(shape=gamma_shape_parameter, scale=1., size=ndims))
    if max_eigvalue is not None:
      eigenvalues *= max_eigvalue / eigenvalues.max()
    q, r = np.linalg.qr(rng.randn(ndims, ndims))
    q *= np.sign(np.diag(r))
    self.covariance = (q * eigenvalues).dot(q.T)
    self.covariance_eigenvalues = eigenvalues
    super().__init__(covariance=self.covariance)


class LogisticRegression(Target):
  """weights ~ N(0, I); labels ~ Bernoulli(logits = [features, 1] @ weights).

  `tensor_core_transitions=True`: HamiltonianMonteCarlo on this target advances ALL chains in lock-step with every
  leapfrog's log-prob + gradient evaluated as two tcgen05 contractions around the sigmoid (pb2_logistic_tc_leapfrog;
  a fixed L makes lock-step free of waste) instead of the FP32 warp-per-chain kernel -- the faster path from a few
  tens of thousands of chains (measured 2.2x at 65,536 chains); trajectories agree to float32 rounding."""
  kind = _lib.TARGET_LOGISTIC

  def __init__(self, train_features, train_labels, tensor_core_transitions=False):
    self.is_lockstep = bool(tensor_core_transitions)
    X = np.asarray(train_features, np.float32)
    X = np.concatenate([X, np.ones([X.shape[0], 1], np.float32)], axis=-1)   # logistic_regression.py:36-39
    y = np.asarray(train_labels).astype(np.float32)
    if y.shape != (X.shape[0],):
      raise ValueError('train_labels must have shape [num_train_points]')
    self.features_with_bias = X
    self.labels = y
    if self.is_lockstep and X.shape[1] > 32:
      raise ValueError('tensor_core_transitions: at most 32 weights (incl. bias); use RowShardedLogisticRegression')
    super().__init__(dim=X.shape[1], n_rows=X.shape[0], a=X, b=y)

  def log_prob_and_grad_tc(self, x):
    """The same value + gradient for all chains at once on the tensor cores (pb2_logistic_logp_grad_tc)."""
    import torch
    x = x.contiguous().float()
    ctx = _lib.Context.get(x.device)
    ctx.bind_stream()
    lp = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    g = torch.empty_like(x)
    _lib.check(ctx.lib.pb2_logistic_logp_grad_tc(ctx.handle, self.handle(ctx), x.shape[0], _lib.ptr(x), _lib.ptr(lp),
                                                 _lib.ptr(g)), ctx.handle)
    return lp, g

  def leapfrog(self, m, x, lp, g, step, step_kind, num_steps):
    """Lock-step SimpleLeapfrogIntegrator on the tensor cores (tensor_core_transitions=True): one C-ABI call."""
    import torch
    ctx = _lib.Context.get(x.device)
    ctx.bind_stream()
    m_out, x_out, g_out = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    lp_out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    _lib.check(ctx.lib.pb2_logistic_tc_leapfrog(
        ctx.handle, self.handle(ctx), x.shape[0], _lib.ptr(m.contiguous()), _lib.ptr(x.contiguous()),
        _lib.ptr(lp.contiguous()), _lib.ptr(g.contiguous()), _lib.ptr(step), step_kind, int(num_steps),
        _lib.ptr(m_out), _lib.ptr(x_out), _lib.ptr(lp_out), _lib.ptr(g_out)), ctx.handle)
    return m_out, x_out, lp_out, g_out


class StochasticVolatility(Target):
  """State (unconstrained): [logit-ish persistence, mean_log_volatility, softplus^-1 shock scale,
  std_log_volatility[T]]."""
  kind = _lib.TARGET_STOCH_VOL

  def __init__(self, centered_returns):
    y = np.asarray(centered_returns, np.float32)
    self.centered_returns = y
    super().__init__(dim=y.size + 3, n_rows=y.size, a=y, part_sizes=[1, 1, 1, y.size])

  @staticmethod
  def constrain(u):
    """default_event_space_bijector forward (vectorized_stochastic_volatility.py:346-356):
    Sigmoid(-1, 1), Identity, Softplus, Identity -- elementwise torch ops on the result."""
    import torch
    out = u.clone()
    out[..., 0] = 2.0 * torch.sigmoid(u[..., 0]) - 1.0
    out[..., 2] = torch.nn.functional.softplus(u[..., 2])
    return out


class StochasticVolatilityConstrained(Target):
  """The same model in its own (constrained) coordinates [persistence_of_volatility in (-1, 1), mean_log_volatility,
  white_noise_shock_scale > 0, std_log_volatility[T]] -- what inference_gym's VectorizedStochasticVolatility.log_prob
  takes (vectorized_stochastic_volatility.py:233-309).  Sample it through
  `TransformedTransitionKernel(kernel, StochasticVolatilityConstrained.default_event_space_bijector())`."""
  kind = _lib.TARGET_STOCH_VOL_CONSTRAINED

  def __init__(self, centered_returns):
    y = np.asarray(centered_returns, np.float32)
    self.centered_returns = y
    super().__init__(dim=y.size + 3, n_rows=y.size, a=y, part_sizes=[1, 1, 1, y.size])

  @staticmethod
  def default_event_space_bijector():
    """vectorized_stochastic_volatility.py:346-356."""
    from probability_b200 import bijectors as b
    return [b.Sigmoid(-1., 1.), b.Identity(), b.Softplus(), b.Identity()]


class StochasticVolatilityCentered(Target):
  """The CENTRED stochastic-volatility model (inference_gym/targets/stochastic_volatility.py:39-111): one latent
  log-volatility x_t per time step -- x_0 ~ N(m, s / sqrt(1 - phi^2)), x_t ~ N(m + phi (x_{t-1} - m), s),
  y_t ~ N(0, exp(x_t / 2)) -- in unconstrained coordinates [logit-ish persistence, mean_log_volatility,
  softplus^-1 shock scale, log_volatility[T]].  Same posterior as `StochasticVolatility` (the non-centred form), a
  harder geometry; its gradient is a 3-point stencil."""
  kind = _lib.TARGET_STOCH_VOL_CENTERED
  constrain = staticmethod(StochasticVolatility.constrain)

  def __init__(self, centered_returns):
    y = np.asarray(centered_returns, np.float32)
    self.centered_returns = y
    super().__init__(dim=y.size + 3, n_rows=y.size, a=y, part_sizes=[1, 1, 1, y.size])


class StochasticVolatilityCenteredConstrained(StochasticVolatilityCentered):
  """The centred model in its own coordinates [persistence in (-1, 1), mean, shock scale > 0, log_volatility[T]]; sample it
  through `TransformedTransitionKernel(kernel, default_event_space_bijector())`."""
  kind = _lib.TARGET_STOCH_VOL_CENTERED_CONSTRAINED
  default_event_space_bijector = staticmethod(StochasticVolatilityConstrained.default_event_space_bijector)


class RowShardedLogisticRegression(Target):
  """Large-data logistic regression whose ROWS are sharded over the ranks of a
  torch.distributed process group (BASELINE config 5).  Every rank holds all chains
  (replicated, identical RNG: no fold_in_axis_index, like an unsharded state part in
  hmc_test.py:1246-1272) and `features_local`/`labels_local`, its slice of the data.

  One gradient evaluation = local pass over the rows for ALL chains (pb2_rowshard_logistic_grad_tc: two tcgen05
  contractions around the sigmoid; pb2_rowshard_logistic_grad, FP32, below 128 chains)
  + ONE all-reduce of the packed `[B, D+1]` (gradient | log-lik) buffer over NVLink (the psum of
  distribute_lib.py:179-186 and its pbroadcast VJP :228-242) + prior (pb2_rowshard_logistic_finish).
  Transitions with this target run lock-step over chains (one kernel sequence per leapfrog)."""
  kind = None
  is_lockstep = True

  def __init__(self, features_local, labels_local, process_group=None, add_bias=True):
    X = np.asarray(features_local, np.float32)
    if add_bias:
      X = np.concatenate([X, np.ones([X.shape[0], 1], np.float32)], axis=-1)
    y = np.asarray(labels_local).astype(np.float32)
    if y.shape != (X.shape[0],):
      raise ValueError('labels_local must have shape [num_local_points]')
    self.dim = X.shape[1]
    self.n_rows = X.shape[0]
    self.part_sizes = [self.dim]
    if self.dim > 100:
      raise ValueError('row-sharded logistic regression supports at most 100 weights (incl. bias)')
    self.padded_dim = 32 if self.dim <= 32 else (64 if self.dim <= 64 else 100)
    Xp = np.zeros([X.shape[0], self.padded_dim], np.float32)
    Xp[:, :self.dim] = X
    self._Xp = Xp
    self._y = y
    self.process_group = process_group
    self.use_tensor_cores = True    # False: always the FP32 kernel (A/B)
    self._dev = {}

  def _device_data(self, device):
    import torch
    key = device.index
    if key not in self._dev:
      self._dev[key] = (torch.from_numpy(self._Xp).to(device), torch.from_numpy(self._y).to(device))
    return self._dev[key]

  def handle(self, ctx):
    raise _lib.Pb2Error('RowShardedLogisticRegression runs the lock-step path; it has no per-chain kernel')

  def _tc_planes(self, ctx, device):
    """This shard's rows pre-split into the tensor-core operand planes (built once per device)."""
    import torch
    key = ('tc', device.index)
    if key not in self._dev:
      X, _ = self._device_data(device)
      nbytes = int(ctx.lib.pb2_rowshard_tc_planes_bytes(self.n_rows))
      planes = torch.empty(nbytes, dtype=torch.uint8, device=device)
      _lib.check(ctx.lib.pb2_rowshard_tc_prepare(ctx.handle, _lib.ptr(X), self.n_rows, self.dim, self.padded_dim,
                                                 _lib.ptr(planes)), ctx.handle)
      self._dev[key] = planes
    return self._dev[key]

  def _world(self):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
      return dist
    return None

  def log_prob_and_grad(self, x):
    import torch
    if x.dim() != 2 or x.shape[1] != self.dim:
      raise ValueError('expected state of shape [chains, {}], got {}'.format(self.dim, tuple(x.shape)))
    x = x.contiguous().float()
    ctx = _lib.Context.get(x.device)
    ctx.bind_stream()
    X, y = self._device_data(x.device)
    B = x.shape[0]
    packed = torch.empty(B, self.dim + 1, dtype=torch.float32, device=x.device)
    # tcgen05 contractions for a tile's worth of chains or more; the FP32 thread-per-chain kernel for small batches
    if B >= 128 and self.use_tensor_cores:
      planes = self._tc_planes(ctx, x.device)
      _lib.check(ctx.lib.pb2_rowshard_logistic_grad_tc(ctx.handle, _lib.ptr(planes), _lib.ptr(y), self.n_rows, self.dim,
                                                       _lib.ptr(x), B, _lib.ptr(packed)), ctx.handle)
    else:
      _lib.check(ctx.lib.pb2_rowshard_logistic_grad(ctx.handle, _lib.ptr(X), _lib.ptr(y), self.n_rows, self.dim,
                                                    self.padded_dim, _lib.ptr(x), B, _lib.ptr(packed)), ctx.handle)
    dist = self._world()
    if dist is not None:
      if self._library_owns_collective(ctx):
        _lib.check(ctx.lib.pb2_comm_allreduce_sum(ctx.handle, _lib.ptr(packed), packed.numel()), ctx.handle)
      else:
        dist.all_reduce(packed, group=self.process_group)     # per-leapfrog gradient all-reduce (NCCL/NVLink)
    lp = torch.empty(B, dtype=torch.float32, device=x.device)
    g = torch.empty_like(x)
    _lib.check(ctx.lib.pb2_rowshard_logistic_finish(ctx.handle, _lib.ptr(packed), _lib.ptr(x), B, self.dim,
                                                    _lib.ptr(g), _lib.ptr(lp)), ctx.handle)
    return lp, g

  def _library_owns_collective(self, ctx):
    """True when the per-leapfrog all-reduce can be issued by libpb2 itself: single rank, or a communicator of the
    group's size attached to the context (distribute.init_comm)."""
    dist = self._world()
    if dist is None:
      return True
    return int(ctx.lib.pb2_comm_size(ctx.handle)) == dist.get_world_size(self.process_group)

  def leapfrog(self, m, x, lp, g, step, step_kind, num_steps):
    """SimpleLeapfrogIntegrator (leapfrog_integrator.py:280-309) lock-step over all chains.  ONE C-ABI call
    (pb2_rowshard_leapfrog) enqueues all leapfrogs -- gradient kernel, NCCL all-reduce, fused prior + kick + drift --
    when the library owns the collective; otherwise the per-leapfrog loop below with torch.distributed."""
    import torch
    ctx = _lib.Context.get(x.device)
    ctx.bind_stream()
    B, D = x.shape
    if self._library_owns_collective(ctx):
      X, y = self._device_data(x.device)
      use_tc = B >= 128 and self.use_tensor_cores
      planes = self._tc_planes(ctx, x.device) if use_tc else None
      m_out, x_out, g_out = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
      lp_out = torch.empty(B, dtype=torch.float32, device=x.device)
      _lib.check(ctx.lib.pb2_rowshard_leapfrog(
          ctx.handle, _lib.ptr(planes), _lib.ptr(X), _lib.ptr(y), self.n_rows, self.dim, self.padded_dim, B,
          _lib.ptr(m.contiguous()), _lib.ptr(x.contiguous()), _lib.ptr(lp.contiguous()), _lib.ptr(g.contiguous()),
          _lib.ptr(step), step_kind, int(num_steps), 0 if self._world() is None else 1, _lib.ptr(m_out),
          _lib.ptr(x_out), _lib.ptr(lp_out),
          _lib.ptr(g_out)), ctx.handle)
      return m_out, x_out, lp_out, g_out
    x = x.clone()
    v = torch.empty_like(x)
    m_out = torch.empty_like(x)
    call = lambda mode, gg: _lib.check(ctx.lib.pb2_lockstep_leapfrog(
        ctx.handle, mode, B, D, _lib.ptr(step), step_kind, _lib.ptr(v), _lib.ptr(x), _lib.ptr(gg), _lib.ptr(m),
        _lib.ptr(m_out)), ctx.handle)
    call(0, g.contiguous())
    for i in range(int(num_steps)):
      lp, g = self.log_prob_and_grad(x)
      call(1 if i + 1 < int(num_steps) else 2, g)
    return m_out, x, lp, g


def synthetic_sv_returns(T=2516, phi=0.95, s=0.25, m=None, seed=0):
  rng = np.random.default_rng(seed)
  m = 2.0 * np.log(15.0) if m is None else m
  h = np.empty(T)
  h[0] = s * rng.standard_normal() / np.sqrt(1 - phi * phi)
  for t in range(1, T):
    h[t] = phi * h[t - 1] + s * rng.standard_normal()
  y = rng.standard_normal(T) * np.exp(0.5 * (h + m))
  return (y - y.mean()).astype(np.float32)


def synthetic_logistic_data(n=1000, d=24, seed=0):
  """Features WITHOUT the bias column (LogisticRegression adds it) and 0/1 labels."""
  rng = np.random.default_rng(seed)
  X = rng.standard_normal((n, d))
  X = (X - X.mean(0)) / X.std(0)
  Xb = np.concatenate([X, np.ones((n, 1))], axis=1)
  theta = rng.standard_normal(d + 1)
  p = 1.0 / (1.0 + np.exp(-(Xb @ theta)))
  y = (rng.random(n) < p).astype(np.float32)
  return X.astype(np.float32), y
