"""DualAveragingStepSizeAdaptation (tfp/mcmc/dual_averaging_step_size_adaptation.py:74-644).

The cross-chain `reduce_logmeanexp` of `min(0, log_accept_ratio)` (math/generic.py:221-274)
and the Nesterov dual-averaging update (:419-475) run as device kernels (pb2_da_partial /
pb2_da_apply); with chains sharded over ranks the (max, sum-exp) partials are all-gathered
through torch.distributed (NCCL) -- the analogue of experimental_reduce_chain_axis_names
(:259-261, distribute_lib.reduce_logsumexp :147-162).
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
      raise NotImplementedError('custom reduce_fn is not supported: the chain reduction is the fused '
                                'log-mean-exp kernel (the reference default)')

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

  def bootstrap_results(self, init_state):
    import torch
    inner_results = self.inner_kernel.bootstrap_results(init_state)
    step_size = self.step_size_getter_fn(inner_results)
    if _engine.is_list_like(step_size):
      if len(step_size) != 1:
        raise NotImplementedError('per-part step-size lists are not supported by the dual-averaging '
                                  'kernel; use one scalar step size')
      step_size = step_size[0]
    dev = step_size.device
    f = lambda v: torch.as_tensor(v, dtype=torch.float32, device=dev)
    st = self._parameters['shrinkage_target']
    log_shrink = (float(np.log(10.)) + torch.log(step_size)) if st is None else torch.log(f(st))
    return DualAveragingStepSizeAdaptationResults(
        inner_results=inner_results, step=torch.tensor(0, dtype=torch.int32, device=dev),
        target_accept_prob=f(self._parameters['target_accept_prob']), log_shrinkage_target=log_shrink,
        exploration_shrinkage=f(self._parameters['exploration_shrinkage']),
        step_count_smoothing=f(self._parameters['step_count_smoothing']),
        decay_rate=f(self._parameters['decay_rate']), error_sum=torch.zeros_like(step_size),
        log_averaging_step=torch.zeros_like(step_size), new_step_size=step_size,
        num_adaptation_steps=torch.tensor(int(self.num_adaptation_steps), dtype=torch.int32, device=dev))

  # -- device state vector (layout in include/pb2.h) -------------------------
  @staticmethod
  def _pack(r):
    import torch
    if r.new_step_size.numel() != 1:
      raise NotImplementedError('dual averaging supports one chain-shared scalar step size')
    z = torch.zeros((), dtype=torch.float32, device=r.new_step_size.device)
    vals = [r.error_sum.reshape(()), r.log_averaging_step.reshape(()), r.log_shrinkage_target.reshape(()),
            r.step.float(), r.num_adaptation_steps.float(), r.target_accept_prob, r.exploration_shrinkage,
            r.step_count_smoothing, r.decay_rate, r.new_step_size.reshape(())] + [z] * 6
    return torch.stack([v.float() for v in vals]).contiguous()

  @staticmethod
  def _unpack(r, st, inner_results):
    shp = r.new_step_size.shape
    return r._replace(inner_results=inner_results, error_sum=st[0].reshape(shp).clone(),
                      log_averaging_step=st[1].reshape(shp).clone(), step=st[3].round().int(),
                      new_step_size=st[9].reshape(shp).clone())

  def _world(self):
    if not self.experimental_reduce_chain_axis_names:
      return None
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
      return dist
    return None

  def one_step(self, current_state, previous_kernel_results, seed=None):
    import torch
    pkr = previous_kernel_results
    inner_results = self.step_size_setter_fn(pkr.inner_results, pkr.new_step_size)        # :482-485
    new_state, new_inner_results = self.inner_kernel.one_step(current_state, inner_results, seed=seed)
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
      cnt = torch.tensor([float(lar.numel())], device=lar.device)
      dist.all_reduce(cnt)
      n_global = int(cnt.item())
      partial = gathered.contiguous()
    n_part = partial.numel() // 2
    _lib.check(ctx.lib.pb2_da_apply(ctx.handle, _lib.ptr(partial), n_part, n_global, _lib.ptr(st), None),
               ctx.handle)
    return new_state, self._unpack(pkr, st, new_inner_results)
