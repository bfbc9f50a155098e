"""SimpleStepSizeAdaptation (tfp/mcmc/simple_step_size_adaptation.py:91-482): multiply or
divide the step size by (1 + adaptation_rate) according to the sign of
(log mean accept prob - log target).  The chain reduction reuses the dual-averaging
log-mean-exp kernel (pb2_da_partial)."""
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

  inner_kernel = property(lambda self: self._parameters['inner_kernel'])
  num_adaptation_steps = property(lambda self: self._parameters['num_adaptation_steps'])
  name = property(lambda self: self._parameters['name'])

  @property
  def is_calibrated(self):
    return self.inner_kernel.is_calibrated

  def bootstrap_results(self, init_state):
    import torch
    inner_results = self.inner_kernel.bootstrap_results(init_state)
    step_size = self._parameters['step_size_getter_fn'](inner_results)
    if _engine.is_list_like(step_size):
      step_size = step_size[0]
    dev = step_size.device
    f = lambda v: torch.as_tensor(v, dtype=torch.float32, device=dev)
    return SimpleStepSizeAdaptationResults(
        inner_results=inner_results, step=torch.tensor(0, dtype=torch.int32, device=dev),
        target_accept_prob=f(self._parameters['target_accept_prob']),
        adaptation_rate=f(self._parameters['adaptation_rate']), new_step_size=step_size,
        num_adaptation_steps=torch.tensor(int(self.num_adaptation_steps), dtype=torch.int32, device=dev))

  def one_step(self, current_state, previous_kernel_results, seed=None):
    import torch
    pkr = previous_kernel_results
    inner_results = self._parameters['step_size_setter_fn'](pkr.inner_results, pkr.new_step_size)
    new_state, new_inner = self.inner_kernel.one_step(current_state, inner_results, seed=seed)
    lar = da._innermost(new_inner).log_accept_ratio.contiguous().float()
    ctx = _lib.Context.get(lar.device)
    ctx.bind_stream()
    partial = torch.empty(2, dtype=torch.float32, device=lar.device)
    _lib.check(ctx.lib.pb2_da_partial(ctx.handle, _lib.ptr(lar), lar.numel(), _lib.ptr(partial)), ctx.handle)
    log_mean = partial[0] + torch.log(partial[1]) - float(np.log(lar.numel()))
    step_size = self._parameters['step_size_getter_fn'](new_inner)
    if _engine.is_list_like(step_size):
      step_size = step_size[0]
    one_plus = 1. + pkr.adaptation_rate                                   # :419-426
    adapted = torch.where(log_mean > torch.log(pkr.target_accept_prob), step_size * one_plus,
                          step_size / one_plus)
    new_step = torch.where(pkr.step < pkr.num_adaptation_steps, adapted, step_size)
    return new_state, pkr._replace(inner_results=new_inner, step=pkr.step + 1, new_step_size=new_step)
