"""DualAveragingStepSizeAdaptation (tfp/mcmc/dual_averaging_step_size_adaptation.py:74-644).

Two execution paths, same arithmetic (:419-475):

* the common case -- ONE chain-shared scalar step size, per-chain `log_accept_ratio [chains]` on the GPU and the
  default getters -- runs the cross-chain `reduce_logmeanexp` of `min(0, log_accept_ratio)`
  (math/generic.py:221-274) and the Nesterov dual-averaging update as device kernels (pb2_da_partial /
  pb2_da_apply); `sample_chain` fuses this case into pb2_run (one C-ABI crossing for the whole run, with the
  cross-rank reduction done inside the library when a communicator is attached, see pb2_comm_init);
* every other reference-legal form -- per-part step-size lists, per-chain `[chains, 1]` step sizes, step sizes that
  broadcast against event dimensions, custom getters -- runs `_one_step_part` literally (:353-475) as a handful
  of torch ops on the (tiny) step-size tensors, on whatever device they live on.

With chains sharded over ranks (`experimental_reduce_chain_axis_names`) the reduction over the chain axis spans
the ranks: the analogue of distribute_lib.reduce_logsumexp (:147-162) over torch.distributed (NCCL).
"""
import collections

import numpy as np

from probability_b200 import _lib
from probability_b200.mcmc import _engine
from probability_b200.mcmc import kernel as kernel_base

DualAveragingStepSizeAdaptationResults = collections.namedtuple(
    'DualAveragingStepSizeAdaptationResults',
    ['inner_results', 'target_accept_prob', 'log_shrinkage_target', 'exploration_shrinkage',
     'step_count_smoothing', 'decay_rate', 'error_sum', 'log_averaging_step', 'step', 'new_step_size',
     'num_adaptation_steps'])


# simple_step_size_adaptation.py:32-48 ---------------------------------------------------
def _innermost(kernel_results):
  kr = kernel_results
  while hasattr(kr, 'inner_results'):
    kr = kr.inner_results
  return kr


def hmc_like_step_size_getter_fn(kernel_results):
  kr = _innermost(kernel_results)
  if hasattr(kr, 'accepted_results'):
    kr = kr.accepted_results
  return kr.step_size


def hmc_like_step_size_setter_fn(kernel_results, new_step_size):
  def rec(kr):
    if hasattr(kr, 'inner_results'):
      return kr._replace(inner_results=rec(kr.inner_results))
    if hasattr(kr, 'accepted_results'):
      return kr._replace(accepted_results=kr.accepted_results._replace(step_size=new_step_size))
    return kr._replace(step_size=new_step_size)
  return rec(kernel_results)


def hmc_like_log_accept_prob_getter_fn(kernel_results):
  import torch
  lar = _innermost(kernel_results).log_accept_ratio
  safe = torch.where(torch.isfinite(lar), lar, torch.full_like(lar, -np.inf))
  return torch.minimum(safe, torch.zeros_like(safe))


def _enable_store_parameters(kernel):
  """mcmc/internal/util.py enable_store_parameters_in_results."""
  k = kernel
  if 'store_parameters_in_results' in k.parameters and not k.parameters['store_parameters_in_results']:
    return k.copy(store_parameters_in_results=True)
  if 'inner_kernel' in k.parameters:
    return k.copy(inner_kernel=_enable_store_parameters(k.parameters['inner_kernel']))
  return k


# ---- chain reductions (math/generic.py:221-274, distribute_lib.py:147-162) ---------------------------------
def get_differing_dims(a, b):
  """Indices of the dimensions where the shapes of `a` and `b` differ; `a` may have fewer dimensions
  (simple_step_size_adaptation.py:51-60)."""
  a_shape = np.int32(tuple(a.shape))
  b_shape = np.int32(tuple(b.shape))
  return [int(i) for i in np.where(a_shape != b_shape[:len(a_shape)])[0]]


def world_of(axis_names, group=None):
  """torch.distributed if chains are sharded over > 1 ranks and a named chain axis is given, else None."""
  if not axis_names:
    return None
  import torch.distributed as dist
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
    return dist
  return None


def reduce_logmeanexp(x, axis, keepdims=False, dist=None, group=None):
  """log(mean(exp(x))) over `axis` (list of dims); with `dist`, the mean also spans the ranks (every rank holds a
  slice of the leading chain axis): global max, then a sum of exp(x - max) and of the element count."""
  import torch
  axis = [int(a) for a in axis]
  if not axis:
    return x
  n = 1
  for a in axis:
    n *= x.shape[a]
  m = torch.amax(x, dim=axis, keepdim=True)
  cnt = torch.tensor(float(n), dtype=x.dtype, device=x.device)
  if dist is not None:
    m = m.contiguous()
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(cnt, group=group)
  m_safe = torch.where(torch.isfinite(m), m, torch.zeros_like(m))          # generic.py:250-256 (all -inf)
  s = torch.sum(torch.exp(x - m_safe), dim=axis, keepdim=True)
  if dist is not None:
    s = s.contiguous()
    dist.all_reduce(s, group=group)
  out = torch.log(s) + m_safe - torch.log(cnt)
  if not keepdims:
    out = out.reshape([d for i, d in enumerate(out.shape) if i not in axis])
  return out


def _flat(s):
  return list(s) if _engine.is_list_like(s) else [s]


def _pack_as(like, parts):
  return type(like)(parts) if _engine.is_list_like(like) else parts[0]


def _as_f32(v, device):
  import torch
  return torch.as_tensor(v, dtype=torch.float32, device=device)


class DualAveragingStepSizeAdaptation(kernel_base.TransitionKernel):

  def __init__(self, inner_kernel, num_adaptation_steps, target_accept_prob=0.75, exploration_shrinkage=0.05,
               shrinkage_target=None, step_count_smoothing=10, decay_rate=0.75,
               step_size_setter_fn=hmc_like_step_size_setter_fn,
               step_size_getter_fn=hmc_like_step_size_getter_fn,
               log_accept_prob_getter_fn=hmc_like_log_accept_prob_getter_fn, reduce_fn=None,
               experimental_reduce_chain_axis_names=None, validate_args=False, name=None):
    inner_kernel = _enable_store_parameters(inner_kernel)
    self._parameters = dict(
        inner_kernel=inner_kernel, num_adaptation_steps=num_adaptation_steps,
        target_accept_prob=target_accept_prob, exploration_shrinkage=exploration_shrinkage,
        shrinkage_target=shrinkage_target, step_count_smoothing=step_count_smoothing, decay_rate=decay_rate,
        step_size_setter_fn=step_size_setter_fn, step_size_getter_fn=step_size_getter_fn,
        log_accept_prob_getter_fn=log_accept_prob_getter_fn, reduce_fn=reduce_fn,
        experimental_reduce_chain_axis_names=experimental_reduce_chain_axis_names,
        validate_args=validate_args, name=name)
    if reduce_fn is not None:
      raise NotImplementedError('custom reduce_fn is not supported: the chain reduction is the '
                                'log-mean-exp of the reference default (reduce_logmeanexp)')
    if validate_args:
      t = np.asarray(target_accept_prob if not hasattr(target_accept_prob, 'cpu')
                     else target_accept_prob.cpu().numpy())
      if np.any(t <= 0.):
        raise ValueError('`target_accept_prob` must be > 0.')
      if np.any(t >= 1.):
        raise ValueError('`target_accept_prob` must be < 1.')

  inner_kernel = property(lambda self: self._parameters['inner_kernel'])
  num_adaptation_steps = property(lambda self: self._parameters['num_adaptation_steps'])
  name = property(lambda self: self._parameters['name'])
  experimental_reduce_chain_axis_names = property(
      lambda self: self._parameters['experimental_reduce_chain_axis_names'])

  def step_size_setter_fn(self, kernel_results, new_step_size):
    return self._parameters['step_size_setter_fn'](kernel_results, new_step_size)

  def step_size_getter_fn(self, kernel_results):
    return self._parameters['step_size_getter_fn'](kernel_results)

  def log_accept_prob_getter_fn(self, kernel_results):
    return self._parameters['log_accept_prob_getter_fn'](kernel_results)

  @property
  def is_calibrated(self):
    return self.inner_kernel.is_calibrated

  @property
  def experimental_shard_axis_names(self):
    return self.inner_kernel.experimental_shard_axis_names

  def experimental_with_shard_axes(self, shard_axis_names):
    return self.copy(inner_kernel=self.inner_kernel.experimental_with_shard_axes(shard_axis_names))

  def _world(self):
    return world_of(self.experimental_reduce_chain_axis_names)

  # ---- bootstrap (:543-609) ----------------------------------------------------------------------------------
  def bootstrap_results(self, init_state):
    inner_results = self.inner_kernel.bootstrap_results(init_state)
    return self._bootstrap_from_inner_results(init_state, inner_results)

  def _bootstrap_from_inner_results(self, init_state, inner_results):
    import torch
    step_size = self.step_size_getter_fn(inner_results)
    log_accept_prob = self.log_accept_prob_getter_fn(inner_results)
    dev = log_accept_prob.device
    state_parts = [_as_f32(s, dev) for s in _flat(init_state)]
    step_size_parts = [_as_f32(s, dev) for s in _flat(step_size)]
    st = self._parameters['shrinkage_target']
    if st is None:
      shrink_parts = [None] * len(step_size_parts)
    else:
      shrink_parts = _flat(st)
      if len(shrink_parts) not in [1, len(step_size_parts)]:
        raise ValueError('`shrinkage_target` should be a Tensor or list of tensors of same length as '
                         '`step_size`. Found len(`step_size`) = {} and len(shrinkage_target) = {}'.format(
                             len(step_size_parts), len(shrink_parts)))
      if len(shrink_parts) < len(step_size_parts):
        shrink_parts = shrink_parts * len(step_size_parts)
    error_sum, log_averaging_step, log_shrinkage_target = [], [], []
    for state_part, step_size_part, shrink in zip(state_parts, step_size_parts, shrink_parts):
      n_red = min(log_accept_prob.dim(), state_part.dim() - step_size_part.dim())
      red = reduce_logmeanexp(log_accept_prob, list(range(max(n_red, 0))))      # shapes only: no collective
      red = reduce_logmeanexp(red, get_differing_dims(red, step_size_part), keepdims=True)
      error_sum.append(torch.zeros_like(red, dtype=torch.float32))
      log_averaging_step.append(torch.zeros_like(step_size_part))
      if shrink is None:
        log_shrinkage_target.append(float(np.log(10.)) + torch.log(step_size_part))
      else:
        log_shrinkage_target.append(torch.log(_as_f32(shrink, dev)))
    f = lambda v: _as_f32(v, dev)
    return DualAveragingStepSizeAdaptationResults(
        inner_results=inner_results, step=torch.tensor(0, dtype=torch.int32, device=dev),
        target_accept_prob=f(self._parameters['target_accept_prob']), log_shrinkage_target=log_shrinkage_target,
        exploration_shrinkage=f(self._parameters['exploration_shrinkage']),
        step_count_smoothing=f(self._parameters['step_count_smoothing']),
        decay_rate=f(self._parameters['decay_rate']), error_sum=error_sum,
        log_averaging_step=log_averaging_step, new_step_size=step_size,
        num_adaptation_steps=torch.tensor(int(self.num_adaptation_steps), dtype=torch.int32, device=dev))

  # ---- device state vector of the fused scalar path (layout in include/pb2.h) ---------------------------------
  def _is_scalar_case(self, pkr, log_accept_prob=None):
    """One chain-shared scalar step size, default getters, per-chain accept ratios on the GPU."""
    parts = _flat(pkr.new_step_size)
    if len(parts) != 1 or parts[0].numel() != 1 or not parts[0].is_cuda:
      return False
    if pkr.target_accept_prob.numel() != 1:
      return False
    if self._parameters['log_accept_prob_getter_fn'] is not hmc_like_log_accept_prob_getter_fn:
      return False
    if log_accept_prob is not None and (log_accept_prob.dim() != 1 or not log_accept_prob.is_cuda):
      return False
    return True

  @staticmethod
  def _pack(r):
    import torch
    s = _flat(r.new_step_size)[0]
    z = torch.zeros((), dtype=torch.float32, device=s.device)
    one = lambda v: _flat(v)[0].reshape(()).float()
    vals = [one(r.error_sum), one(r.log_averaging_step), one(r.log_shrinkage_target),
            r.step.float(), r.num_adaptation_steps.float(), r.target_accept_prob.reshape(()),
            r.exploration_shrinkage, r.step_count_smoothing, r.decay_rate, s.reshape(()).float()] + [z] * 6
    return torch.stack([v.float() for v in vals]).contiguous()

  @staticmethod
  def _unpack(r, st, inner_results):
    like = lambda old, v: _pack_as(old, [v.reshape(_flat(old)[0].shape).clone()])
    return r._replace(inner_results=inner_results, error_sum=like(r.error_sum, st[0]),
                      log_averaging_step=like(r.log_averaging_step, st[1]), step=st[3].round().int(),
                      new_step_size=like(r.new_step_size, st[9]))

  # ---- one step (:477-532) ---------------------------------------------------------------------------------
  def one_step(self, current_state, previous_kernel_results, seed=None):
    pkr = previous_kernel_results
    inner_results = self.step_size_setter_fn(pkr.inner_results, pkr.new_step_size)        # :482-485
    new_state, new_inner_results = self.inner_kernel.one_step(current_state, inner_results, seed=seed)
    log_accept_prob = self.log_accept_prob_getter_fn(new_inner_results)
    if self._is_scalar_case(pkr, log_accept_prob):
      return new_state, self._one_step_device(pkr, new_inner_results)
    return new_state, self._one_step_general(current_state, pkr, new_inner_results, log_accept_prob)

  def _one_step_device(self, pkr, new_inner_results):
    import torch
    lar = _innermost(new_inner_results).log_accept_ratio.contiguous().float()
    ctx = _lib.Context.get(lar.device)
    ctx.bind_stream()
    st = self._pack(pkr)
    partial = torch.empty(2, dtype=torch.float32, device=lar.device)
    _lib.check(ctx.lib.pb2_da_partial(ctx.handle, _lib.ptr(lar), lar.numel(), _lib.ptr(partial)), ctx.handle)
    dist = self._world()
    n_global = lar.numel()
    if dist is not None:
      ws = dist.get_world_size()
      gathered = torch.empty(ws, 2, dtype=torch.float32, device=lar.device)
      dist.all_gather_into_tensor(gathered, partial)
      shard = getattr(self.inner_kernel, 'chain_shard', None)
      if shard is not None:
        n_global = int(shard.num_chains_global)      # known on the host: no device round trip
      else:
        cnt = torch.tensor([float(lar.numel())], device=lar.device)
        dist.all_reduce(cnt)
        n_global = int(cnt.item())
      partial = gathered.contiguous()
    n_part = partial.numel() // 2
    _lib.check(ctx.lib.pb2_da_apply(ctx.handle, _lib.ptr(partial), n_part, n_global, _lib.ptr(st), None),
               ctx.handle)
    return self._unpack(pkr, st, new_inner_results)

  def _one_step_general(self, current_state, pkr, new_inner_results, log_accept_prob):
    dev = log_accept_prob.device
    step_size = self.step_size_getter_fn(new_inner_results)
    step_size_parts = [_as_f32(s, dev) for s in _flat(step_size)]
    state_parts = [_as_f32(s, dev) for s in _flat(current_state)][:len(step_size_parts)]
    if len(state_parts) < len(step_size_parts):
      raise ValueError('There should be exactly one `step_size` or it should have same length as '
                       '`current_state`.')
    error_sum_parts = _flat(pkr.error_sum)
    las_parts = _flat(pkr.log_averaging_step)
    lst_parts = _flat(pkr.log_shrinkage_target)
    outs = [self._one_step_part(s, x, e, la, ls, log_accept_prob, pkr)
            for s, x, e, la, ls in zip(step_size_parts, state_parts, error_sum_parts, las_parts, lst_parts)]
    new_step, new_las, new_err = zip(*outs)
    return pkr._replace(inner_results=new_inner_results, error_sum=_pack_as(pkr.error_sum, list(new_err)),
                        step=pkr.step + 1, log_averaging_step=_pack_as(pkr.log_averaging_step, list(new_las)),
                        new_step_size=_pack_as(step_size, list(new_step)))

  def _one_step_part(self, step_size, state, error_sum, log_averaging_step, log_shrinkage_target,
                     log_accept_prob, pkr):
    """:353-475, op for op."""
    import torch
    num_reduce_dims = max(min(log_accept_prob.dim(), state.dim() - step_size.dim()), 0)
    # the leading reduced axis is the chain axis: with sharded chains its mean spans the ranks
    dist = self._world() if num_reduce_dims > 0 else None
    red = reduce_logmeanexp(log_accept_prob, list(range(num_reduce_dims)), dist=dist)
    red = reduce_logmeanexp(red, get_differing_dims(red, step_size), keepdims=True)
    new_error_sum = error_sum + pkr.target_accept_prob - torch.exp(red)
    pad = max(log_shrinkage_target.dim() - new_error_sum.dim(), 0)
    ext = new_error_sum.reshape(tuple(new_error_sum.shape) + (1,) * pad)
    step = pkr.step.to(pkr.step_count_smoothing.dtype) + 1.
    soft_t = pkr.step_count_smoothing + step
    new_log_step = log_shrinkage_target - (ext * torch.sqrt(step)) / (soft_t * pkr.exploration_shrinkage)
    eta = step ** (-pkr.decay_rate)
    new_las = eta * new_log_step + (1. - eta) * log_averaging_step
    istep = pkr.step + 1
    n_adapt = pkr.num_adaptation_steps
    new_step_size = torch.where(istep < n_adapt, torch.exp(new_log_step),
                                torch.where(istep > n_adapt, step_size + torch.zeros_like(new_log_step),
                                            torch.exp(new_las)))
    new_las = torch.where(istep > n_adapt, log_averaging_step + torch.zeros_like(new_las), new_las)
    return new_step_size, new_las, new_error_sum
