"""Streaming reducers and `sample_fold` (tfp/experimental/mcmc/{reducer.py, expectations_reducer.py,
covariance_reducer.py:38-220, potential_scale_reduction_reducer.py:36-160, sample_fold.py:35-180, with_reductions.py:41}):
statistics of a chain without materialising its `[num_steps, chains, ...]` history.

The reference folds one state per step into every reducer.  Here `sample_fold` advances the chain in bounded CHUNKS of
fused transitions (one pb2_run per chunk) and every chunk of states `[n, chains, D]` is folded into running per-chain
moments by ONE device pass (pb2_running_moments_update over the `[n, chains * D]` view: Welford / Chan, fixed order), so
the memory in flight is the chunk, not the run.  Every reducer below is a function of those running moments.
"""
import collections

import numpy as np

from probability_b200 import _lib
from probability_b200 import random as pb_random
from probability_b200.mcmc import _engine
from probability_b200.mcmc import kernel as kernel_lib
from probability_b200.mcmc import sample as sample_lib


class _ChainMoments(object):
  """count, mean[B, D], sum of squared deviations[B, D] over the steps seen so far (per chain and dimension)."""

  def __init__(self, B, D, device):
    import torch
    self.B, self.D = B, D
    self.state = torch.zeros(1 + 2 * B * D, dtype=torch.float32, device=device)
    self.comoment = None

  def _update_comoment(self, states):
    """Chan's merge of the chunk's centred Gram matrices (one batched library GEMM per chunk) into the running
    co-moment, using the running count / mean BEFORE this chunk (sample_stats.py RunningCovariance.update)."""
    import torch
    n = float(states.shape[0])
    na = self.count.clone()
    mb = states.mean(0)                                         # [B, D]
    xc = (states - mb).permute(1, 2, 0).contiguous()            # [B, D, n]
    gram = torch.bmm(xc, xc.transpose(1, 2))                    # [B, D, D]
    delta = mb - self.mean
    w = na * n / (na + n)
    self.comoment += gram + w * delta[:, :, None] * delta[:, None, :]

  def track_comoment(self):
    """Also keep the per-chain co-moment matrix sum_t (x_t - mean)(x_t - mean)^T [B, D, D] (CovarianceReducer)."""
    import torch
    if self.comoment is None:
      self.comoment = torch.zeros(self.B, self.D, self.D, dtype=torch.float32, device=self.state.device)

  def update(self, states):
    """states: [n, B, D] float32 CUDA."""
    if self.comoment is not None:
      self._update_comoment(states)
    x = states.reshape(states.shape[0], self.B * self.D).contiguous()
    ctx = _lib.Context.get(x.device)
    ctx.bind_stream()
    _lib.check(ctx.lib.pb2_running_moments_update(ctx.handle, _lib.ptr(x), x.shape[0], x.shape[1], _lib.ptr(self.state)),
               ctx.handle)

  @property
  def count(self):
    return self.state[0]

  @property
  def mean(self):
    return self.state[1:1 + self.B * self.D].reshape(self.B, self.D)

  @property
  def m2(self):
    return self.state[1 + self.B * self.D:].reshape(self.B, self.D)


class Reducer(object):
  """reducer.py:27-110: `initialize(initial_chain_state, initial_kernel_results)`, `one_step(new_chain_state,
  current_reducer_state, previous_kernel_results)`, `finalize(final_reducer_state)`.  `sample_fold` feeds these reducers
  the shared running moments; `one_step` is provided for the per-step protocol (a chunk of one state)."""

  def initialize(self, initial_chain_state, initial_kernel_results=None):
    x, shapes, was_list = _engine.flatten_state(initial_chain_state)
    st = _ChainMoments(x.shape[0], x.shape[1], x.device)
    st.shapes, st.was_list = shapes, was_list
    return st

  def one_step(self, new_chain_state, current_reducer_state, previous_kernel_results=None):
    x, _, _ = _engine.flatten_state(new_chain_state)
    current_reducer_state.update(x[None])
    return current_reducer_state

  def finalize(self, final_reducer_state):
    raise NotImplementedError

  @staticmethod
  def _unflat(st, v):
    return _engine.unflatten(v, st.shapes, st.was_list)


class ExpectationsReducer(Reducer):
  """expectations_reducer.py:30-130 with the identity transform: running mean of the state, per chain."""

  def finalize(self, st):
    return self._unflat(st, st.mean)


class VarianceReducer(Reducer):
  """covariance_reducer.py:223-290: running variance of the state over the steps, per chain (`ddof` as there)."""

  def __init__(self, ddof=0):
    self.ddof = ddof

  def finalize(self, st):
    return self._unflat(st, st.m2 / (st.count - float(self.ddof)))


class CovarianceReducer(Reducer):
  """covariance_reducer.py:43-220: running covariance of the state over the steps, per chain.  `event_ndims` = 1 (every
  state part `[chains, d]`: result `[chains, d, d]` per part) or 0 (variances, the VarianceReducer).  The reference's
  default `event_ndims=None` treats the chain axis as part of the event (a `[chains, d, chains, d]` matrix) and is not
  offered; nor are `transform_fn`s (the reducers see states only)."""
  needs_comoment = True

  def __init__(self, event_ndims=1, ddof=0, name=None):
    del name
    if event_ndims not in (0, 1):
      raise NotImplementedError('CovarianceReducer: event_ndims must be 0 or 1 (covariance within a chain)')
    self.event_ndims, self.ddof = event_ndims, ddof
    self.needs_comoment = event_ndims == 1

  def initialize(self, initial_chain_state, initial_kernel_results=None):
    st = super(CovarianceReducer, self).initialize(initial_chain_state, initial_kernel_results)
    if self.event_ndims == 1:
      if any(len(sh) != 1 for sh in st.shapes):
        raise ValueError('CovarianceReducer(event_ndims=1) needs state parts of shape [chains, d]')
      st.track_comoment()
    return st

  def finalize(self, st):
    denom = st.count - float(self.ddof)
    if self.event_ndims == 0:
      return self._unflat(st, st.m2 / denom)
    if st.comoment is None:
      raise ValueError('the reducer state does not carry co-moments (initialize it with this reducer)')
    cov = st.comoment / denom
    outs, off = [], 0
    for n in _engine.part_sizes_of(st.shapes):
      outs.append(cov[:, off:off + n, off:off + n])
      off += n
    return outs if st.was_list else outs[0]


class PotentialScaleReductionReducer(Reducer):
  """potential_scale_reduction_reducer.py:36-160: R-hat from per-chain running means and variances
  (diagnostic.py:476-567 with independent_chain_ndims = 1)."""

  def __init__(self, independent_chain_ndims=1):
    if independent_chain_ndims != 1:
      raise NotImplementedError('one chain axis')

  def finalize(self, st):
    n, m = st.count, float(st.B)
    if st.B < 2:
      raise ValueError('Must provide at least 2 chains.')
    chain_var = st.m2 / (n - 1.0)                       # within-chain variances (ddof = 1)
    w = chain_var.mean(0)
    b_div_n = st.mean.var(0, unbiased=True)             # variance of the chain means
    rhat = ((m + 1.0) / m) * (((n - 1.0) / n) * w + b_div_n) / w - (n - 1.0) / (m * n)
    return _unflat_event(st, rhat)


def _unflat_event(st, v):
  """[D] -> the state's structure without the chain axis."""
  sizes = _engine.part_sizes_of(st.shapes)
  outs, off = [], 0
  for s, n in zip(st.shapes, sizes):
    outs.append(v[off:off + n].reshape(s))
    off += n
  return outs if st.was_list else outs[0]


SampleFoldResults = collections.namedtuple('SampleFoldResults', ['reduction_results', 'end_state',
                                                                 'final_kernel_results'])


def sample_fold(num_steps, current_state, previous_kernel_results=None, kernel=None, reducer=None,
                previous_reducer_state=None, return_final_reducer_states=False, num_burnin_steps=0,
                num_steps_between_results=0, parallel_iterations=10, seed=None, name=None,
                experimental_chunk_bytes=256 << 20):
  """sample_fold.py:35-180: `num_steps` results folded into `reducer` (a Reducer or a list of them); returns
  `(reduction_results, end_state, final_kernel_results)`.  The chain advances in chunks of at most
  `experimental_chunk_bytes` of states; chunk seeds are derived by `seed_chunk, seed = split_seed(seed)`."""
  del parallel_iterations, name
  import torch
  if kernel is None:
    raise ValueError('`kernel` is required')
  reducers = list(reducer) if isinstance(reducer, (list, tuple)) else [reducer]
  single = not isinstance(reducer, (list, tuple))
  seed = pb_random.sanitize_seed(seed, salt='mcmc.sample_fold')
  if previous_kernel_results is None:
    previous_kernel_results = kernel.bootstrap_results(current_state)
  x, shapes, was_list = _engine.flatten_state(current_state)
  B, D = x.shape
  shared = previous_reducer_state
  if shared is None:
    shared = _ChainMoments(B, D, x.device)
    shared.shapes, shared.was_list = shapes, was_list
    if any(getattr(r, 'needs_comoment', False) for r in reducers if r is not None):
      if any(len(sh) != 1 for sh in shapes):
        raise ValueError('CovarianceReducer(event_ndims=1) needs state parts of shape [chains, d]')
      shared.track_comoment()
  chunk = int(max(1, min(int(num_steps), experimental_chunk_bytes // max(1, 4 * B * D))))
  state, pkr = current_state, previous_kernel_results
  done = 0
  while done < int(num_steps):
    n = min(chunk, int(num_steps) - done)
    cseed, seed = pb_random.split_seed(seed)
    res = sample_lib.sample_chain(n, state, previous_kernel_results=pkr, kernel=kernel, trace_fn=None,
                                  num_burnin_steps=num_burnin_steps if done == 0 else num_steps_between_results,
                                  num_steps_between_results=num_steps_between_results,
                                  return_final_kernel_results=True, seed=cseed)
    st = res.all_states
    flat = torch.cat([s.reshape(s.shape[0], s.shape[1], -1) for s in st], -1) if was_list else st.reshape(n, B, D)
    shared.update(flat)
    state = [s[-1] for s in st] if was_list else st[-1]
    pkr = res.final_kernel_results
    done += n
    del res, st, flat
  results = [r.finalize(shared) if r is not None else None for r in reducers]
  out = results[0] if single else results
  if return_final_reducer_states:
    out = (out, shared)
  return SampleFoldResults(out, state, pkr)


WithReductionsKernelResults = collections.namedtuple('WithReductionsKernelResults',
                                                     ['reduction_results', 'inner_results'])


def _map_reducers(fn, reducer, *rest):
  if isinstance(reducer, (list, tuple)):
    return type(reducer)(fn(*args) for args in zip(reducer, *rest))
  return fn(reducer, *rest)


class WithReductions(kernel_lib.TransitionKernel):
  """with_reductions.py:41-180: a TransitionKernel that steps `inner_kernel` and folds every new state into `reducer`
  (a Reducer or a list / tuple of them); the reducer states ride in `kernel_results.reduction_results`, finalize them
  with `reducer.finalize(state)`.  The per-step protocol: one transition + one moments update per `one_step` (use
  `sample_fold` for the chunked, fused form).  Reducer states are updated in place (device buffers), so results of an
  earlier step alias the current ones."""

  def __init__(self, inner_kernel, reducer, adjust_kr_fn=lambda kr: kr, name=None):
    self._parameters = dict(inner_kernel=inner_kernel, reducer=reducer, adjust_kr_fn=adjust_kr_fn, name=name)

  inner_kernel = property(lambda self: self._parameters['inner_kernel'])
  reducer = property(lambda self: self._parameters['reducer'])
  adjust_kr_fn = property(lambda self: self._parameters['adjust_kr_fn'])
  name = property(lambda self: self._parameters['name'])
  parameters = property(lambda self: self._parameters)
  is_calibrated = property(lambda self: self.inner_kernel.is_calibrated)

  def bootstrap_results(self, init_state, inner_results=None, previous_reducer_state=None):
    """The initial state does not count as a sample: the reducer states start as empty streams."""
    if inner_results is None:
      inner_results = self.inner_kernel.bootstrap_results(init_state)
    if previous_reducer_state is None:
      previous_reducer_state = _map_reducers(lambda r: r.initialize(init_state, inner_results), self.reducer)
    return WithReductionsKernelResults(previous_reducer_state, inner_results)

  def one_step(self, current_state, previous_kernel_results, seed=None):
    new_state, inner_results = self.inner_kernel.one_step(current_state, previous_kernel_results.inner_results,
                                                          seed=seed)
    adj = self.adjust_kr_fn(inner_results)
    red = _map_reducers(lambda r, st: r.one_step(new_state, st, previous_kernel_results=adj), self.reducer,
                        previous_kernel_results.reduction_results)
    return new_state, WithReductionsKernelResults(red, inner_results)


def step_kernel(num_steps, current_state, previous_kernel_results=None, kernel=None,
                return_final_kernel_results=False, parallel_iterations=10, seed=None, name=None):
  """step.py:29-106: `num_steps` transitions of `kernel` from `current_state`, nothing traced; returns the end state
  (and the final kernel results).  Seeds follow the reference's loop: `step_seed, seed = split_seed(seed)` per step."""
  del parallel_iterations, name
  if kernel is None:
    raise ValueError('`kernel` is required')
  seed = pb_random.sanitize_seed(seed, salt='mcmc_step_kernel')
  if previous_kernel_results is None:
    previous_kernel_results = kernel.bootstrap_results(current_state)
  state, kr = current_state, previous_kernel_results
  for _ in range(int(num_steps)):
    step_seed, seed = pb_random.split_seed(seed)
    state, kr = kernel.one_step(state, kr, seed=step_seed)
  return (state, kr) if return_final_kernel_results else state


SampleDiscardingKernelResults = collections.namedtuple('SampleDiscardingKernelResults',
                                                       ['call_counter', 'inner_results'])


class SampleDiscardingKernel(kernel_lib.TransitionKernel):
  """sample_discarding_kernel.py:40-175: burn-in and thinning as a kernel -- the first `one_step` advances the inner
  kernel `num_burnin_steps + num_steps_between_results + 1` times, every later one `num_steps_between_results + 1`
  times (through `step_kernel`, seeded with this step's seed), so a `WithReductions` wrapped around it only sees the
  states that survive."""

  def __init__(self, inner_kernel, num_burnin_steps=0, num_steps_between_results=0, name=None):
    self._parameters = dict(inner_kernel=inner_kernel, num_burnin_steps=num_burnin_steps,
                            num_steps_between_results=num_steps_between_results, name=name)

  inner_kernel = property(lambda self: self._parameters['inner_kernel'])
  num_burnin_steps = property(lambda self: self._parameters['num_burnin_steps'])
  num_steps_between_results = property(lambda self: self._parameters['num_steps_between_results'])
  name = property(lambda self: self._parameters['name'])
  is_calibrated = property(lambda self: self.inner_kernel.is_calibrated)

  def bootstrap_results(self, init_state, inner_results=None):
    if inner_results is None:
      inner_results = self.inner_kernel.bootstrap_results(init_state)
    return SampleDiscardingKernelResults(0, inner_results)

  def one_step(self, current_state, previous_kernel_results, seed=None):
    calls = int(previous_kernel_results.call_counter)
    skip = int(self.num_steps_between_results) + (int(self.num_burnin_steps) if calls == 0 else 0)
    state, inner = step_kernel(skip + 1, current_state, previous_kernel_results=previous_kernel_results.inner_results,
                               kernel=self.inner_kernel, return_final_kernel_results=True, seed=seed)
    return state, SampleDiscardingKernelResults(calls + 1, inner)
