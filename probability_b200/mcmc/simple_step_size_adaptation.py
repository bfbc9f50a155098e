"""SimpleStepSizeAdaptation (tfp/mcmc/simple_step_size_adaptation.py:91-482): multiply or divide each step-size part
by (1 + adaptation_rate) according to the sign of (log mean accept prob - log target), for the first
`num_adaptation_steps` steps.

The chain reduction is the reference's `reduce_logmeanexp` (the default `reduce_fn`): for one chain-shared scalar
step size on the GPU it is the dual-averaging reduction kernel (pb2_da_partial); every other reference-legal
step-size form (per-part lists, per-chain `[chains, 1]`, event-broadcast) follows :353-440 literally as torch
ops on the step-size tensors.  With chains sharded over ranks (`experimental_reduce_chain_axis_names`) the mean
spans the ranks, exactly like DualAveragingStepSizeAdaptation."""
import collections

import numpy as np

from probability_b200 import _lib
from probability_b200.mcmc import _engine
from probability_b200.mcmc import dual_averaging_step_size_adaptation as da
from probability_b200.mcmc import kernel as kernel_base

SimpleStepSizeAdaptationResults = collections.namedtuple(
    'SimpleStepSizeAdaptationResults',
    ['inner_results', 'target_accept_prob', 'adaptation_rate', 'step', 'new_step_size',
     'num_adaptation_steps'])

hmc_like_step_size_getter_fn = da.hmc_like_step_size_getter_fn
hmc_like_step_size_setter_fn = da.hmc_like_step_size_setter_fn
hmc_like_log_accept_prob_getter_fn = da.hmc_like_log_accept_prob_getter_fn
get_differing_dims = da.get_differing_dims


class SimpleStepSizeAdaptation(kernel_base.TransitionKernel):

  def __init__(self, inner_kernel, num_adaptation_steps, target_accept_prob=0.75, adaptation_rate=0.01,
               step_size_setter_fn=hmc_like_step_size_setter_fn,
               step_size_getter_fn=hmc_like_step_size_getter_fn,
               log_accept_prob_getter_fn=hmc_like_log_accept_prob_getter_fn, reduce_fn=None,
               experimental_reduce_chain_axis_names=None, validate_args=False, name=None):
    inner_kernel = da._enable_store_parameters(inner_kernel)
    self._parameters = dict(
        inner_kernel=inner_kernel, num_adaptation_steps=num_adaptation_steps,
        target_accept_prob=target_accept_prob, adaptation_rate=adaptation_rate,
        step_size_setter_fn=step_size_setter_fn, step_size_getter_fn=step_size_getter_fn,
        log_accept_prob_getter_fn=log_accept_prob_getter_fn, reduce_fn=reduce_fn,
        experimental_reduce_chain_axis_names=experimental_reduce_chain_axis_names,
        validate_args=validate_args, name=name)
    if reduce_fn is not None:
      raise NotImplementedError('custom reduce_fn is not supported: the chain reduction is the '
                                'log-mean-exp of the reference default (reduce_logmeanexp)')

  inner_kernel = property(lambda self: self._parameters['inner_kernel'])
  num_adaptation_steps = property(lambda self: self._parameters['num_adaptation_steps'])
  name = property(lambda self: self._parameters['name'])
  experimental_reduce_chain_axis_names = property(
      lambda self: self._parameters['experimental_reduce_chain_axis_names'])

  @property
  def is_calibrated(self):
    return self.inner_kernel.is_calibrated

  @property
  def experimental_shard_axis_names(self):
    return self.inner_kernel.experimental_shard_axis_names

  def experimental_with_shard_axes(self, shard_axis_names):
    return self.copy(inner_kernel=self.inner_kernel.experimental_with_shard_axes(shard_axis_names))

  def _world(self):
    return da.world_of(self.experimental_reduce_chain_axis_names)

  def bootstrap_results(self, init_state):
    import torch
    inner_results = self.inner_kernel.bootstrap_results(init_state)
    step_size = self._parameters['step_size_getter_fn'](inner_results)
    dev = da._flat(step_size)[0].device if torch.is_tensor(da._flat(step_size)[0]) else None
    if dev is None:
      dev = self._parameters['log_accept_prob_getter_fn'](inner_results).device
      step_size = da._pack_as(step_size, [da._as_f32(s, dev) for s in da._flat(step_size)])
    f = lambda v: da._as_f32(v, dev)
    return SimpleStepSizeAdaptationResults(
        inner_results=inner_results, step=torch.tensor(0, dtype=torch.int32, device=dev),
        target_accept_prob=f(self._parameters['target_accept_prob']),
        adaptation_rate=f(self._parameters['adaptation_rate']), new_step_size=step_size,
        num_adaptation_steps=torch.tensor(int(self.num_adaptation_steps), dtype=torch.int32, device=dev))

  def _log_mean_device(self, lar):
    """log-mean-exp of min(0, finite_or(-inf)(log_accept_ratio)) over all chains (and ranks): pb2_da_partial."""
    import torch
    lar = lar.contiguous().float()
    ctx = _lib.Context.get(lar.device)
    ctx.bind_stream()
    partial = torch.empty(2, dtype=torch.float32, device=lar.device)
    _lib.check(ctx.lib.pb2_da_partial(ctx.handle, _lib.ptr(lar), lar.numel(), _lib.ptr(partial)), ctx.handle)
    # the partial is a 64-bit fixed-point sum (2^-36 units) of the accept probabilities: integer sums over the ranks
    # are exact, so every sharding of the chains adapts with the same bits
    total = partial.view(torch.int64).clone()
    cnt = torch.tensor([lar.numel()], dtype=torch.int64, device=lar.device)
    dist = self._world()
    if dist is not None:
      dist.all_reduce(total)
      dist.all_reduce(cnt)
    mean = total.to(torch.float64) / (2.0 ** 36) / cnt.to(torch.float64)
    return torch.log(mean).to(torch.float32).reshape(())

  def one_step(self, current_state, previous_kernel_results, seed=None):
    import torch
    pkr = previous_kernel_results
    inner_results = self._parameters['step_size_setter_fn'](pkr.inner_results, pkr.new_step_size)   # :361-365
    new_state, new_inner = self.inner_kernel.one_step(current_state, inner_results, seed=seed)
    getter = self._parameters['log_accept_prob_getter_fn']
    log_accept_prob = getter(new_inner)
    dev = log_accept_prob.device
    log_target = torch.log(pkr.target_accept_prob.to(log_accept_prob.dtype))
    step_size = self._parameters['step_size_getter_fn'](new_inner)
    step_size_parts = [da._as_f32(s, dev) for s in da._flat(step_size)]
    state_parts = [da._as_f32(s, dev) for s in da._flat(current_state)]
    scalar_case = (len(da._flat(step_size)) == 1 and step_size_parts[0].numel() == 1 and
                   log_accept_prob.dim() == 1 and log_accept_prob.is_cuda and
                   getter is hmc_like_log_accept_prob_getter_fn)
    one_plus = 1. + pkr.adaptation_rate
    new_parts = []
    for step_size_part, state_part in zip(step_size_parts, state_parts):
      if scalar_case:
        red = self._log_mean_device(da._innermost(new_inner).log_accept_ratio).reshape(step_size_part.shape)
      else:
        n_red = max(min(log_accept_prob.dim(), state_part.dim() - step_size_part.dim()), 0)       # :407-409
        dist = self._world() if n_red > 0 else None
        red = da.reduce_logmeanexp(log_accept_prob, list(range(n_red)), dist=dist)
        red = da.reduce_logmeanexp(red, get_differing_dims(red, step_size_part), keepdims=True)
      adapted = torch.where(red > log_target, step_size_part * one_plus, step_size_part / one_plus)   # :419-426
      new_parts.append(torch.where(pkr.step < pkr.num_adaptation_steps, adapted,
                                   step_size_part + torch.zeros_like(adapted)))
    return new_state, pkr._replace(inner_results=new_inner, step=pkr.step + 1,
                                   new_step_size=da._pack_as(step_size, new_parts))
