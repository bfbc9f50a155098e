"""sample_chain (tfp/mcmc/sample.py:81-383; loop semantics internal/loop_util.py:69-253).

When the kernel is one of the fused CUDA kernels (HamiltonianMonteCarlo, NoUTurnSampler,
optionally wrapped in DualAveragingStepSizeAdaptation) and `trace_fn` only selects
kernel-results fields, the whole call is ONE crossing of the C ABI (pb2_run): seeds are
chained `step_seed, seed = split_seed(seed)` per transition on the host, every transition
of every chain runs inside persistent kernels, and traced fields are written straight to
their `[num_results, chains, ...]` output tensors.  Anything else runs the reference's
step-by-step loop over `kernel.one_step` (still CUDA transitions, one launch per step).
"""
import collections
import warnings

import numpy as np

from probability_b200 import _lib
from probability_b200 import random as pb_random
from probability_b200.mcmc import _engine
from probability_b200.mcmc import dual_averaging_step_size_adaptation as da_lib
from probability_b200.mcmc import hmc as hmc_lib
from probability_b200.mcmc import nuts as nuts_lib

StatesAndTrace = collections.namedtuple('StatesAndTrace', 'all_states, trace')
CheckpointableStatesAndTrace = collections.namedtuple(
    'CheckpointableStatesAndTrace', 'all_states, trace, final_kernel_results')


def _default_trace_fn(current_state, kernel_results):
  return kernel_results


class _Spy(object):
  """Stands in for the kernel results while probing which fields `trace_fn` selects."""

  def __init__(self, path=()):
    object.__setattr__(self, '_path', path)

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    return _Spy(self._path + (name,))


def _map_structure(fn, s):
  if isinstance(s, _Spy):
    return fn(s)
  if hasattr(s, '_fields'):
    return type(s)(*[_map_structure(fn, v) for v in s])
  if isinstance(s, (list, tuple)):
    return type(s)(_map_structure(fn, v) for v in s)
  if isinstance(s, dict):
    return type(s)((k, _map_structure(fn, v)) for k, v in s.items())
  return fn(s)


def _stack_structure(items):
  import torch
  first = items[0]
  if hasattr(first, '_fields'):
    return type(first)(*[_stack_structure([it[i] for it in items]) for i in range(len(first))])
  if isinstance(first, (list, tuple)):
    return type(first)(_stack_structure([it[i] for it in items]) for i in range(len(first)))
  if isinstance(first, dict):
    return type(first)((k, _stack_structure([it[k] for it in items])) for k in first)
  if torch.is_tensor(first):
    return torch.stack(items)
  if isinstance(first, np.ndarray):
    return np.stack(items)
  return np.asarray(items)


def _probe_trace_fn(trace_fn, state_struct):
  """Returns (structure with _Spy leaves, set of paths) or None if trace_fn computes on values."""
  try:
    out = trace_fn(state_struct, _Spy())
  except Exception:  # pylint: disable=broad-except
    return None
  paths = []
  ok = [True]

  def visit(leaf):
    if isinstance(leaf, _Spy):
      paths.append(leaf._path)
    else:
      ok[0] = False
    return leaf

  _map_structure(visit, out)
  if not ok[0]:
    return None
  return out, paths


def _fused_kernel(kernel):
  """-> (inner fused kernel, dual-averaging wrapper or None) if the stack is fusable."""
  if isinstance(kernel, (hmc_lib.HamiltonianMonteCarlo, nuts_lib.NoUTurnSampler)):
    return kernel, None
  if isinstance(kernel, da_lib.DualAveragingStepSizeAdaptation) and isinstance(
      kernel.inner_kernel, (hmc_lib.HamiltonianMonteCarlo, nuts_lib.NoUTurnSampler)):
    if kernel._world() is not None:
      # chains sharded over ranks: fused only when the library owns the collective (distribute.init_comm);
      # otherwise the step loop with torch.distributed collectives
      from probability_b200 import distribute
      try:
        if distribute.comm_size() != kernel._world().get_world_size():
          return None
      except _lib.Pb2Error:
        return None
    return kernel.inner_kernel, kernel
  return None


_NUTS_ROOT = tuple((f,) for f in nuts_lib.NoUTurnSampler._ALL)


def _try_fused_transformed(kernel, num_results, current_state, pkr, num_burnin_steps, num_steps_between_results,
                           trace_fn, seed, leapfrog_total):
  """TransformedTransitionKernel around a fusable stack: the whole run happens in the unconstrained space (one pb2_run),
  the traced states are mapped forward afterwards and `trace_fn` sees results nested under `inner_results`."""
  from probability_b200.mcmc import transformed_kernel as ttk_lib
  inner_trace = None if trace_fn is None else (
      lambda s, kr: trace_fn(s, ttk_lib.TransformedTransitionKernelResults(transformed_state=None, inner_results=kr)))
  if trace_fn is not None:
    probe = _probe_trace_fn(trace_fn, current_state)
    if probe is None or any((not q) or q[0] != 'inner_results' for q in probe[1]):
      return None
  out = _try_fused(kernel.inner_kernel, num_results, pkr.transformed_state, pkr.inner_results, num_burnin_steps,
                   num_steps_between_results, inner_trace, seed, leapfrog_total)
  if out is None:
    return None
  t_states, trace, final_inner, seed_out = out
  was_list = _engine.is_list_like(t_states)
  last = [s[-1] for s in t_states] if was_list else t_states[-1]
  final = ttk_lib.TransformedTransitionKernelResults(transformed_state=last, inner_results=final_inner)
  return kernel._forward(t_states), trace, final, seed_out


def _try_fused(kernel, num_results, current_state, pkr, num_burnin_steps, num_steps_between_results,
               trace_fn, seed, leapfrog_total=None):
  from probability_b200.mcmc import transformed_kernel as ttk_lib
  if isinstance(kernel, ttk_lib.TransformedTransitionKernel):
    return _try_fused_transformed(kernel, num_results, current_state, pkr, num_burnin_steps,
                                  num_steps_between_results, trace_fn, seed, leapfrog_total)
  fk = _fused_kernel(kernel)
  if fk is None:
    return None
  inner, da = fk
  x, shapes, was_list = _engine.flatten_state(current_state)
  x = x.clone()
  if trace_fn is None:
    struct, paths = (), []
  else:
    probe = _probe_trace_fn(trace_fn, current_state)
    if probe is None:
      return None
    struct, paths = probe
  # strip the dual-averaging prefix
  inner_paths = []
  expand_root = {}
  for p in paths:
    q = p
    if da is not None:
      if not q or q[0] != 'inner_results':
        return None
      q = q[1:]
    if q == ():
      if not isinstance(inner, nuts_lib.NoUTurnSampler):
        return None
      expand_root[p] = True
      inner_paths.extend(_NUTS_ROOT + (('step_size',),))
    else:
      inner_paths.append(q)
  inner_pkr = pkr.inner_results if da is not None else pkr
  da_state = None
  step = None
  if da is not None:
    if not da._is_scalar_case(pkr):
      return None   # per-part / per-chain step sizes, custom getters: the reference's step loop (general update)
    inner_pkr = da.step_size_setter_fn(inner_pkr, pkr.new_step_size)
    da_state = da._pack(pkr)
    step = da_lib._flat(pkr.new_step_size)[0].reshape(1).float().contiguous().clone()
  res = inner._fused_run(x, shapes, was_list, inner_pkr, seed, num_results, num_burnin_steps,
                         num_steps_between_results, list(dict.fromkeys(inner_paths)), da_state=da_state,
                         step=step, leapfrog_total=leapfrog_total,
                         da_over_ranks=(da is not None and da._world() is not None))
  if res is None:
    return None
  states, traced, final_inner, seed_out = res

  def fill(leaf):
    p = leaf._path
    q = p[1:] if da is not None else p
    if p in expand_root:
      vals = {f: traced[(f,)] for f in nuts_lib.NoUTurnSampler._ALL}
      vals['step_size'] = traced.get(('step_size',))
      vals['seed'] = None
      return nuts_lib.NUTSKernelResults(**vals)
    return traced[q]

  trace = _map_structure(fill, struct) if trace_fn is not None else ()
  final = final_inner
  if da is not None:
    final = da._unpack(pkr, da_state, final_inner)
  all_states = _engine.unflatten(states, shapes, was_list)
  return all_states, trace, final, seed_out


def sample_chain(num_results, current_state, previous_kernel_results=None, kernel=None, num_burnin_steps=0,
                 num_steps_between_results=0, trace_fn=_default_trace_fn, return_final_kernel_results=False,
                 parallel_iterations=10, seed=None, name=None, experimental_leapfrog_total=None):
  """Markov chain sampling: same signature and return structure as tfp.mcmc.sample_chain.

  Traced fields come back stacked along a leading `num_results` axis.  `trace_fn` is applied to
  the stacked kernel results when it only selects fields (fused path) and per step otherwise.
  `experimental_leapfrog_total`: optional uint64 CUDA tensor [chains] accumulating gradient
  evaluations (fused path only).
  """
  del parallel_iterations, name
  if kernel is None:
    raise ValueError('`kernel` is required')
  if seed is None:
    seed = np.random.default_rng().integers(0, 2**32, size=2, dtype=np.uint32)
  seed = pb_random.sanitize_seed(seed, salt='mcmc.sample_chain')          # sample.py:312
  if not kernel.is_calibrated:
    warnings.warn('supplied `TransitionKernel` is not calibrated. Markov chain may not converge to '
                  'intended target distribution.')
  num_results = int(num_results)
  num_burnin_steps = int(num_burnin_steps)
  num_steps_between_results = int(num_steps_between_results)
  if previous_kernel_results is None:
    previous_kernel_results = kernel.bootstrap_results(current_state)
  no_trace = trace_fn is None
  if trace_fn is _default_trace_fn:
    warnings.warn('Tracing all kernel results by default is deprecated. Set the `trace_fn` argument to '
                  'None (the future default value) or an explicit callback that traces the values you '
                  'are interested in.')

  fused = _try_fused(kernel, num_results, current_state, previous_kernel_results, num_burnin_steps,
                     num_steps_between_results, trace_fn, seed, experimental_leapfrog_total)
  if fused is not None:
    all_states, trace, final_kernel_results, _ = fused
  else:
    state, pkr = current_state, previous_kernel_results
    states, traces = [], []
    for r in range(num_results):
      n = 1 + (num_burnin_steps if r == 0 else num_steps_between_results)
      for _ in range(n):
        step_seed, seed = pb_random.split_seed(seed)                       # sample.py:344-349
        state, pkr = kernel.one_step(state, pkr, seed=step_seed)
      states.append(state)
      traces.append(() if no_trace else trace_fn(state, pkr))
    all_states = _stack_structure(states)
    trace = () if no_trace else _stack_structure(traces)
    final_kernel_results = pkr

  if return_final_kernel_results:
    return CheckpointableStatesAndTrace(all_states=all_states, trace=trace,
                                        final_kernel_results=final_kernel_results)
  if no_trace:
    return all_states
  return StatesAndTrace(all_states=all_states, trace=trace)
