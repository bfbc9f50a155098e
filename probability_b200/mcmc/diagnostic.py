"""effective_sample_size / potential_scale_reduction (tfp/mcmc/diagnostic.py:38-567).

Both run as CUDA reductions over `states[N, chains, D]` (pb2_ess: direct lagged
auto-covariance with early termination instead of the reference's zero-padded FFT,
stats/sample_stats.py:44-215; pb2_rhat: per-dimension cross-chain reduction).
"""
import numpy as np

from probability_b200 import _lib


def _prep(states):
  """-> (x [N,B,D] float32 contiguous CUDA, batch_shape, event_shape)."""
  import torch
  if not torch.is_tensor(states):
    states = torch.as_tensor(np.asarray(states))
  if not states.is_cuda:
    if not torch.cuda.is_available():
      raise _lib.Pb2Error('probability_b200 diagnostics need a CUDA device; there is no CPU fallback.')
    states = states.cuda()
  if not states.is_floating_point():
    states = states.float()   # diagnostic.py:480-485 casts integer states to float
  return states.float().contiguous()


def _ess_single(states, filter_threshold, filter_beyond_lag, filter_beyond_positive_pairs, cross_chain_dims):
  import torch
  x = _prep(states)
  if x.dim() < 1:
    raise ValueError('states must have a leading sample dimension')
  N = x.shape[0]
  if cross_chain_dims is not None:
    ccd = cross_chain_dims if cross_chain_dims >= 0 else cross_chain_dims + x.dim()
    if ccd != 1:
      raise NotImplementedError('only cross_chain_dims=1 (states [N, chains, ...]) is supported')
    if x.dim() < 2 or x.shape[1] < 2:
      raise ValueError('When `cross_chain_dims` is not `None`, there must be > 1 chain in `states`.')
    B = x.shape[1]
    rest = tuple(x.shape[2:])
    out_shape = rest
  else:
    B = 1
    rest = tuple(x.shape[1:])
    out_shape = rest
  D = int(np.prod(rest)) if rest else 1
  x3 = x.reshape(N, B, D)
  ctx = _lib.Context.get(x.device)
  ctx.bind_stream()
  out = torch.empty(D if cross_chain_dims is not None else B * D, dtype=torch.float32, device=x.device)
  thr = float('nan') if filter_threshold is None else float(filter_threshold)
  lag = -1 if filter_beyond_lag is None else int(filter_beyond_lag)
  _lib.check(ctx.lib.pb2_ess(ctx.handle, _lib.ptr(x3), N, B, D, thr, lag,
                             1 if filter_beyond_positive_pairs else 0,
                             1 if cross_chain_dims is not None else 0, _lib.ptr(out)), ctx.handle)
  return out.reshape(out_shape)


def effective_sample_size(states, filter_threshold=0., filter_beyond_lag=None,
                          filter_beyond_positive_pairs=False, cross_chain_dims=None, validate_args=False,
                          name=None):
  """Same arguments as tfp.mcmc.effective_sample_size; `states` is a tensor `[N, ...]` or a
  list of such tensors (the other arguments then broadcast over the list)."""
  del validate_args, name
  if isinstance(states, (list, tuple)):
    n = len(states)

    def bc(v):
      return list(v) if isinstance(v, (list, tuple)) else [v] * n
    return type(states)(
        _ess_single(s, ft, fl, fp, cc) for s, ft, fl, fp, cc in zip(
            states, bc(filter_threshold), bc(filter_beyond_lag), bc(filter_beyond_positive_pairs),
            bc(cross_chain_dims)))
  return _ess_single(states, filter_threshold, filter_beyond_lag, filter_beyond_positive_pairs,
                     cross_chain_dims)


def _rhat_single(state, independent_chain_ndims, split_chains):
  import torch
  x = _prep(state)
  if independent_chain_ndims < 1:
    raise ValueError('Argument `independent_chain_ndims` must be `>= 1`, found: {}'.format(
        independent_chain_ndims))
  N = x.shape[0]
  if split_chains and N < 4:
    raise ValueError('Must provide at least 4 samples when splitting chains. Found {}'.format(N))
  if not split_chains and N < 2:
    raise ValueError('Must provide at least 2 samples.  Found {}'.format(N))
  chain_shape = tuple(x.shape[1:1 + independent_chain_ndims])
  rest = tuple(x.shape[1 + independent_chain_ndims:])
  B = int(np.prod(chain_shape)) if chain_shape else 1
  D = int(np.prod(rest)) if rest else 1
  x3 = x.reshape(N, B, D)
  ctx = _lib.Context.get(x.device)
  ctx.bind_stream()
  out = torch.empty(D, dtype=torch.float32, device=x.device)
  _lib.check(ctx.lib.pb2_rhat(ctx.handle, _lib.ptr(x3), N, B, D, 1 if split_chains else 0, _lib.ptr(out)),
             ctx.handle)
  return out.reshape(rest)


def potential_scale_reduction(chains_states, independent_chain_ndims=1, split_chains=False,
                              validate_args=False, name=None):
  del validate_args, name
  if isinstance(chains_states, (list, tuple)):
    return type(chains_states)(_rhat_single(s, independent_chain_ndims, split_chains) for s in chains_states)
  return _rhat_single(chains_states, independent_chain_ndims, split_chains)
