"""TransformedTransitionKernel (tfp/mcmc/transformed_kernel.py:165-440): runs the inner kernel in the unconstrained
space of `bijector` (one elementwise bijector per state part) and reports states in the constrained space.

The inner kernel's `target_log_prob_fn` is replaced by `TransformedTarget`: the CUDA transition kernels evaluate
log_prob(forward(u)) + forward_log_det_jacobian(u) and its gradient in registers (pb2_targets.cuh TransformedT; the
bijectors cross the C ABI as per-dimension codes, pb2_run_cfg.d_bijector_*), so nothing about the transitions changes.
"""
import collections

import numpy as np

from probability_b200 import _lib
from probability_b200 import bijectors as bij_lib
from probability_b200 import targets as pb_targets
from probability_b200.mcmc import _engine
from probability_b200.mcmc import kernel as kernel_base

TransformedTransitionKernelResults = collections.namedtuple(
    'TransformedTransitionKernelResults', ['transformed_state', 'inner_results'])


class TransformedTarget(pb_targets.Target):
  """`make_transformed_log_prob` (transformed_kernel.py:86-140) for a fused target: same device handle, plus the
  per-dimension bijector arrays that ride along every launch."""

  def __init__(self, base, bijectors):
    self.base = base
    self.kind = base.kind
    self.dim, self.n_rows, self.part_sizes = base.dim, base.n_rows, list(base.part_sizes)
    self.bijectors = list(bijectors)
    if len(self.bijectors) != len(self.part_sizes):
      raise ValueError('need one bijector per state part: the target has {} parts, got {} bijectors'.format(
          len(self.part_sizes), len(self.bijectors)))
    kind, lo, hi = [], [], []
    for b, n in zip(self.bijectors, self.part_sizes):
      kind += [int(b.code)] * n
      lo += [float(b.low)] * n
      hi += [float(b.high)] * n
    self._kind = np.asarray(kind, np.int32)
    self._lo = np.asarray(lo, np.float32)
    self._hi = np.asarray(hi, np.float32)
    self._dev = {}

  def handle(self, ctx):
    return self.base.handle(ctx)

  def bijector_arrays(self, device):
    """(kind int32 [D], low [D], high [D]) on `device`."""
    import torch
    key = str(device)
    if key not in self._dev:
      self._dev[key] = (torch.from_numpy(self._kind).to(device), torch.from_numpy(self._lo).to(device),
                        torch.from_numpy(self._hi).to(device))
    return self._dev[key]

  def log_prob_and_grad(self, x):
    import torch
    if x.dim() != 2 or x.shape[1] != self.dim:
      raise ValueError('expected state of shape [chains, {}], got {}'.format(self.dim, tuple(x.shape)))
    x = x.contiguous().float()
    ctx = _lib.Context.get(x.device)
    ctx.bind_stream()
    k, lo, hi = self.bijector_arrays(x.device)
    lp = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    g = torch.empty_like(x)
    _lib.check(ctx.lib.pb2_logp_grad_transformed(ctx.handle, self.handle(ctx), x.shape[0], _lib.ptr(x), _lib.ptr(k),
                                                 _lib.ptr(lo), _lib.ptr(hi), _lib.ptr(lp), _lib.ptr(g)), ctx.handle)
    return lp, g


def _as_list(x):
  return list(x) if _engine.is_list_like(x) else [x]


class TransformedTransitionKernel(kernel_base.TransitionKernel):

  def __init__(self, inner_kernel, bijector, name=None):
    self._parameters = dict(inner_kernel=inner_kernel, bijector=bijector, name=name)
    self._bijectors = _as_list(bijector)
    for b in self._bijectors:
      if not isinstance(b, bij_lib.Bijector):
        raise TypeError('bijector must be a probability_b200.bijectors.Bijector (Identity, Exp, Softplus, Sigmoid) or a '
                        'list of them, one per state part; got {!r}'.format(b))
    target = inner_kernel.parameters.get('target_log_prob_fn')
    if target is None:
      raise ValueError('inner_kernel must have a target_log_prob_fn (transformed_kernel.py:232-237)')
    self._transformed_target = TransformedTarget(_engine.require_target(target), self._bijectors)
    self._inner_kernel = inner_kernel.copy(target_log_prob_fn=self._transformed_target)

  inner_kernel = property(lambda self: self._inner_kernel)
  bijector = property(lambda self: self._parameters['bijector'])
  name = property(lambda self: self._parameters['name'])

  @property
  def is_calibrated(self):
    return self._inner_kernel.is_calibrated

  @property
  def experimental_shard_axis_names(self):
    return self._inner_kernel.experimental_shard_axis_names

  def experimental_with_shard_axes(self, shard_axis_names):
    return self.copy(inner_kernel=self._parameters['inner_kernel'].experimental_with_shard_axes(shard_axis_names))

  def _forward(self, parts):
    was_list = _engine.is_list_like(parts)
    out = [b.forward(p) for b, p in zip(self._bijectors, _as_list(parts))]
    return out if was_list else out[0]

  def _inverse(self, parts):
    import torch
    was_list = _engine.is_list_like(parts)
    out = [b.inverse(torch.as_tensor(p, dtype=torch.float32)) for b, p in zip(self._bijectors, _as_list(parts))]
    return out if was_list else out[0]

  def bootstrap_results(self, init_state=None, transformed_init_state=None):
    """Exactly one of `init_state` (constrained) / `transformed_init_state` (unconstrained) (:383-440)."""
    if (init_state is None) == (transformed_init_state is None):
      raise ValueError('Must specify exactly one of `init_state` or `transformed_init_state`.')
    if transformed_init_state is None:
      transformed_init_state = self._inverse(init_state)
    return TransformedTransitionKernelResults(
        transformed_state=transformed_init_state,
        inner_results=self._inner_kernel.bootstrap_results(transformed_init_state))

  def one_step(self, current_state, previous_kernel_results, seed=None):
    """`current_state` is ignored in favour of the unconstrained state kept in the results (:343-368)."""
    del current_state
    pkr = previous_kernel_results
    t_next, inner = self._inner_kernel.one_step(pkr.transformed_state, pkr.inner_results, seed=seed)
    return self._forward(t_next), TransformedTransitionKernelResults(transformed_state=t_next, inner_results=inner)
