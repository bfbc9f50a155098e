"""Oracle diagnostics: ESS and R-hat (TEST INFRASTRUCTURE).

  auto_correlation            tfp/stats/sample_stats.py:115-215 (FFT, zero-pad to pow2 >= 2N)
  effective_sample_size       tfp/mcmc/diagnostic.py:203-336
  potential_scale_reduction   tfp/mcmc/diagnostic.py:476-567, _reduce_variance :571-580
"""
import numpy as np


def auto_covariance(x, max_lags=None):
  """x: [N, ...] real; returns [min(N, max_lags+1), ...] unbiased (divide by N-k),
  centred, normalize=False (sample_stats.py:128-213)."""
  x = np.asarray(x)
  dt = x.dtype
  n = x.shape[0]
  xr = np.moveaxis(x, 0, -1)
  xr = xr - xr.mean(axis=-1, keepdims=True)
  target = int(2 ** np.ceil(np.log(n * 2.0) / np.log(2.0)))
  cdt = np.complex64 if dt == np.float32 else np.complex128
  pad = np.zeros(xr.shape[:-1] + (target,), cdt)
  pad[..., :n] = xr
  f = np.fft.fft(pad, axis=-1).astype(cdt)
  sp = np.fft.ifft(f * np.conj(f), axis=-1).astype(cdt)
  sp = sp.real.astype(dt)
  max_lags = n - 1 if max_lags is None else min(n - 1, int(max_lags))
  sp = sp[..., :max_lags + 1]
  denom = (n - np.arange(0.0, max_lags + 1.0)).astype(dt)
  return np.moveaxis(sp / denom, -1, 0)


def reduce_variance(x, axis=None, biased=True, keepdims=False):
  """diagnostic.py:571-580."""
  x = np.asarray(x)
  mean = np.mean(x, axis=axis, keepdims=True)
  bv = np.mean((x - mean) ** 2, axis=axis, keepdims=keepdims)
  if biased:
    return bv
  if axis is None:
    n = x.size
  else:
    n = int(np.prod([x.shape[a] for a in np.atleast_1d(axis)]))
  return (n / (n - 1.0)) * bv


def effective_sample_size(states, filter_threshold=0.0, filter_beyond_lag=None,
                          filter_beyond_positive_pairs=False, cross_chain_dims=None):
  """diagnostic.py:203-336 for one state tensor [N, ...]."""
  states = np.asarray(states)
  dt = states.dtype
  auto_cov = auto_covariance(states, filter_beyond_lag)
  n = dt.type(states.shape[0])
  if cross_chain_dims is not None:
    ccd = cross_chain_dims if cross_chain_dims >= 0 else cross_chain_dims + states.ndim
    num_chains = states.shape[ccd]
    if num_chains < 2:
      raise ValueError('When `cross_chain_dims` is not `None`, there must be > 1 chain in `states`.')
    between = reduce_variance(states.mean(axis=0), biased=False, axis=ccd - 1)
    biased_within = auto_cov[0].mean(axis=ccd - 1)
    approx_var = biased_within + between
    mean_auto_cov = auto_cov.mean(axis=ccd)
    auto_corr = 1.0 - (biased_within - mean_auto_cov) / approx_var
  else:
    auto_corr = auto_cov / auto_cov[:1]
    num_chains = 1
  k = np.arange(0.0, auto_corr.shape[0]).astype(dt)
  nk = ((n - k) / n).reshape([-1] + [1] * (auto_corr.ndim - 1))
  weighted = nk * auto_corr
  if filter_beyond_positive_pairs:
    def sum_pairs(a):
      ln = a.shape[0]
      a = a[:ln - ln % 2]
      return a.reshape((ln // 2, 2) + a.shape[1:]).sum(axis=1)
    mask = (sum_pairs(auto_corr) < 0).astype(dt)
    mask = np.cumsum(mask, axis=0)
    mask = np.maximum(1.0 - mask, 0.0)
    weighted = sum_pairs(weighted) * mask
  elif filter_threshold is not None:
    mask = (auto_corr < filter_threshold).astype(dt)
    mask = np.cumsum(mask, axis=0)
    mask = np.maximum(1.0 - mask, 0.0)
    weighted = weighted * mask
  return (num_chains * n / (-1 + 2 * weighted.sum(axis=0))).astype(dt)


def potential_scale_reduction(state, independent_chain_ndims=1, split_chains=False):
  """diagnostic.py:476-567 for one state tensor [N, chains..., event...]."""
  state = np.asarray(state)
  if state.dtype == np.int64:
    state = state.astype(np.float64)
  elif np.issubdtype(state.dtype, np.integer):
    state = state.astype(np.float32)
  n_samples = state.shape[0]
  if split_chains and n_samples < 4:
    raise ValueError('Must provide at least 4 samples when splitting chains. Found {}'.format(n_samples))
  if not split_chains and n_samples < 2:
    raise ValueError('Must provide at least 2 samples.  Found {}'.format(n_samples))
  if split_chains:
    state = state[:n_samples - n_samples % 2]
    state = state.reshape((2, n_samples // 2) + state.shape[1:])
    state = np.swapaxes(state, 0, 1)
    independent_chain_ndims += 1
  sample_axis = (0,)
  chain_axis = tuple(range(1, 1 + independent_chain_ndims))
  sc_axis = tuple(range(0, 1 + independent_chain_ndims))
  n = float(state.shape[0])
  m = float(np.prod([state.shape[a] for a in chain_axis]))
  b_div_n = reduce_variance(state.mean(axis=sample_axis, keepdims=True), sc_axis, biased=False)
  w = reduce_variance(state, sample_axis, keepdims=True, biased=False).mean(axis=sc_axis)
  sigma2 = ((n - 1) / n) * w + b_div_n
  return (((m + 1.0) / m) * sigma2 / w - (n - 1.0) / (m * n)).astype(state.dtype)
