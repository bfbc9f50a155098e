"""Hamiltonian Monte Carlo kernels (tfp/mcmc/hmc.py, tfp/mcmc/metropolis_hastings.py).

  HamiltonianMonteCarlo             hmc.py:175-540 -- here ONE fused CUDA transition
                                    (momentum draw + L leapfrogs + MH accept) per chain.
  UncalibratedHamiltonianMonteCarlo hmc.py:543-905 -- built from the C-ABI primitives
                                    pb2_rng_normal + pb2_leapfrog (the composable path).
  MetropolisHastings                metropolis_hastings.py:59-302 -- accept/reject with
                                    `choose` (mcmc/internal/util.py:103-164) as torch.where.
"""
import collections

import numpy as np

from probability_b200 import _lib
from probability_b200 import random as pb_random
from probability_b200.mcmc import _engine
from probability_b200.mcmc import kernel as kernel_base

UncalibratedHamiltonianMonteCarloKernelResults = collections.namedtuple(
    'UncalibratedHamiltonianMonteCarloKernelResults',
    ['log_acceptance_correction', 'target_log_prob', 'grads_target_log_prob', 'initial_momentum',
     'final_momentum', 'step_size', 'num_leapfrog_steps', 'seed'])

MetropolisHastingsKernelResults = collections.namedtuple(
    'MetropolisHastingsKernelResults',
    ['accepted_results', 'is_accepted', 'log_accept_ratio', 'proposed_state', 'proposed_results',
     'extra', 'seed'])


def _strip_seed(r):
  return r._replace(seed=[])


def _where(mask, a, b):
  import torch
  if a is b:
    return a
  if isinstance(a, (list, tuple)) and not hasattr(a, '_fields'):
    return type(a)(_where(mask, u, v) for u, v in zip(a, b))
  if not torch.is_tensor(a) or not torch.is_tensor(b):
    return a
  if a.dim() == 0 or a.shape[0] != mask.shape[0]:
    return a
  m = mask.reshape(mask.shape + (1,) * (a.dim() - mask.dim()))
  return torch.where(m, a, b)


def choose(is_accepted, proposed, current):
  """mcmc/internal/util.py:103-164."""
  if hasattr(proposed, '_fields'):
    if not isinstance(proposed, type(current)):
      raise TypeError('Type of `proposed` ({}) must be identical to type of `current` ({}).'.format(
          type(proposed).__name__, type(current).__name__))
    return type(proposed)(**{f: choose(is_accepted, getattr(proposed, f), getattr(current, f))
                             for f in proposed._fields})
  return _where(is_accepted, proposed, current)


class UncalibratedHamiltonianMonteCarlo(kernel_base.TransitionKernel):

  def __init__(self, target_log_prob_fn, step_size, num_leapfrog_steps, state_gradients_are_stopped=False,
               store_parameters_in_results=False, experimental_shard_axis_names=None, name=None):
    if int(num_leapfrog_steps) < 1:
      raise ValueError('num_leapfrog_steps must be >= 1')
    self._parameters = dict(
        target_log_prob_fn=target_log_prob_fn, step_size=step_size, num_leapfrog_steps=num_leapfrog_steps,
        state_gradients_are_stopped=state_gradients_are_stopped,
        store_parameters_in_results=store_parameters_in_results,
        experimental_shard_axis_names=experimental_shard_axis_names, name=name)
    self._target = _engine.require_target(target_log_prob_fn)
    _engine.check_shard_axis_names(experimental_shard_axis_names)

  target_log_prob_fn = property(lambda self: self._parameters['target_log_prob_fn'])
  step_size = property(lambda self: self._parameters['step_size'])
  num_leapfrog_steps = property(lambda self: self._parameters['num_leapfrog_steps'])
  name = property(lambda self: self._parameters['name'])

  @property
  def is_calibrated(self):
    return False

  @property
  def _store_parameters_in_results(self):
    return self._parameters['store_parameters_in_results']

  def bootstrap_results(self, init_state):
    import torch
    x, shapes, was_list = _engine.flatten_state(init_state)
    lp, g = self._target.log_prob_and_grad(x)
    zeros = torch.zeros_like(x)
    res = UncalibratedHamiltonianMonteCarloKernelResults(
        log_acceptance_correction=torch.zeros_like(lp), target_log_prob=lp,
        grads_target_log_prob=_engine.unflatten(g, shapes, True),
        initial_momentum=_engine.unflatten(zeros, shapes, was_list),
        final_momentum=_engine.unflatten(zeros, shapes, was_list),
        step_size=[], num_leapfrog_steps=[], seed=pb_random.zeros_seed())
    if self._store_parameters_in_results:
      res = res._replace(step_size=_as_tensor_struct(self.step_size, x.device),
                         num_leapfrog_steps=torch.tensor(int(self.num_leapfrog_steps), dtype=torch.int32,
                                                         device=x.device))
    return res

  def one_step(self, current_state, previous_kernel_results, seed=None):
    import torch
    pkr = previous_kernel_results
    if self._store_parameters_in_results:
      step_size, L = pkr.step_size, int(pkr.num_leapfrog_steps)
    else:
      step_size, L = self.step_size, int(self.num_leapfrog_steps)
    x, shapes, was_list = _engine.flatten_state(current_state)
    B, D = x.shape
    g, _, _ = _engine.flatten_state(list(pkr.grads_target_log_prob))
    lp = pkr.target_log_prob.contiguous()
    step, step_kind = _engine.step_size_tensor(step_size, B, D, shapes, x.device)
    seed = pb_random.sanitize_seed(seed)
    sizes = _engine.part_sizes_of(shapes)
    seeds = pb_random.split_seed(seed, n=len(sizes))                      # hmc.py:685
    m0 = torch.cat([pb_random.normal((B, n), seed=seeds[i], device=x.device) for i, n in enumerate(sizes)],
                   dim=1).contiguous()                                    # hmc.py:689-695
    if getattr(self._target, 'is_lockstep', False):
      # row-sharded data: all chains advance together, gradient all-reduced every leapfrog
      m1, x1, lp1, g1 = self._target.leapfrog(m0, x, lp, g, step, step_kind, L)
    else:
      ctx = _lib.Context.get(x.device)
      ctx.bind_stream()
      m1, x1, g1 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
      lp1 = torch.empty_like(lp)
      _lib.check(ctx.lib.pb2_leapfrog(ctx.handle, self._target.handle(ctx), B, _lib.ptr(m0), _lib.ptr(x),
                                      _lib.ptr(lp), _lib.ptr(g), _lib.ptr(step), step_kind, L, _lib.ptr(m1),
                                      _lib.ptr(x1), _lib.ptr(lp1), _lib.ptr(g1)), ctx.handle)
    s = (m0 * m0).sum(1) + (-(m1 * m1).sum(1))                            # hmc.py:862-875
    corr = 0.5 * torch.where(torch.isfinite(s), s, torch.full_like(s, -np.inf))
    res = pkr._replace(
        log_acceptance_correction=corr, target_log_prob=lp1,
        grads_target_log_prob=_engine.unflatten(g1, shapes, True),
        initial_momentum=_engine.unflatten(m0, shapes, was_list),
        final_momentum=_engine.unflatten(m1, shapes, was_list), seed=seed)
    return _engine.unflatten(x1, shapes, was_list), res


def _as_tensor_struct(step_size, device):
  import torch
  if _engine.is_list_like(step_size):
    return [torch.as_tensor(s, dtype=torch.float32, device=device) for s in step_size]
  return torch.as_tensor(step_size, dtype=torch.float32, device=device)


class MetropolisHastings(kernel_base.TransitionKernel):
  """Generic MH wrapper over an uncalibrated proposal kernel (metropolis_hastings.py:59)."""

  def __init__(self, inner_kernel, name=None):
    self._parameters = dict(inner_kernel=inner_kernel, name=name)

  inner_kernel = property(lambda self: self._parameters['inner_kernel'])
  name = property(lambda self: self._parameters['name'])

  @property
  def is_calibrated(self):
    return True

  def bootstrap_results(self, init_state):
    import torch
    pkr = self.inner_kernel.bootstrap_results(init_state)
    if not hasattr(pkr, 'target_log_prob'):
      raise ValueError('"target_log_prob" must be a member of `inner_kernel` results.')
    x = pkr.target_log_prob
    return MetropolisHastingsKernelResults(
        accepted_results=_strip_seed(pkr), is_accepted=torch.ones_like(x, dtype=torch.bool),
        log_accept_ratio=torch.zeros_like(x), proposed_state=init_state, proposed_results=pkr, extra=[],
        seed=pb_random.zeros_seed())

  def one_step(self, current_state, previous_kernel_results, seed=None):
    import torch
    seed = pb_random.sanitize_seed(seed)
    proposal_seed, acceptance_seed = pb_random.split_seed(seed)           # metropolis_hastings.py:183
    proposed_state, proposed_results = self.inner_kernel.one_step(
        current_state, previous_kernel_results.accepted_results, seed=proposal_seed)
    if not hasattr(proposed_results, 'target_log_prob'):
      raise ValueError('"target_log_prob" must be a member of `inner_kernel` results.')
    s = proposed_results.target_log_prob + (-previous_kernel_results.accepted_results.target_log_prob)
    s = s + proposed_results.log_acceptance_correction
    log_accept_ratio = torch.where(torch.isfinite(s), s, torch.full_like(s, -np.inf))   # util.py:205-235
    u = pb_random.uniform(tuple(s.shape), seed=acceptance_seed, device=s.device)
    is_accepted = torch.log(u) < log_accept_ratio
    next_state = choose(is_accepted, proposed_state, current_state)
    results = MetropolisHastingsKernelResults(
        accepted_results=choose(is_accepted, _strip_seed(proposed_results),
                                previous_kernel_results.accepted_results),
        is_accepted=is_accepted, log_accept_ratio=log_accept_ratio, proposed_state=proposed_state,
        proposed_results=proposed_results, extra=[], seed=seed)
    return next_state, results


class HamiltonianMonteCarlo(kernel_base.TransitionKernel):
  """Fused HMC transition: MetropolisHastings(UncalibratedHMC) in one persistent kernel."""

  def __init__(self, target_log_prob_fn, step_size, num_leapfrog_steps, state_gradients_are_stopped=False,
               store_parameters_in_results=False, experimental_shard_axis_names=None,
               experimental_chain_shard=None, name=None):
    if int(num_leapfrog_steps) < 1:
      raise ValueError('num_leapfrog_steps must be >= 1')
    self._parameters = dict(
        target_log_prob_fn=target_log_prob_fn, step_size=step_size, num_leapfrog_steps=num_leapfrog_steps,
        state_gradients_are_stopped=state_gradients_are_stopped,
        store_parameters_in_results=store_parameters_in_results,
        experimental_shard_axis_names=experimental_shard_axis_names,
        experimental_chain_shard=experimental_chain_shard, name=name)
    self._target = _engine.require_target(target_log_prob_fn)
    _engine.check_shard_axis_names(experimental_shard_axis_names)
    self._impl = MetropolisHastings(UncalibratedHamiltonianMonteCarlo(
        target_log_prob_fn, step_size, num_leapfrog_steps,
        store_parameters_in_results=store_parameters_in_results))

  target_log_prob_fn = property(lambda self: self._parameters['target_log_prob_fn'])
  step_size = property(lambda self: self._parameters['step_size'])
  num_leapfrog_steps = property(lambda self: self._parameters['num_leapfrog_steps'])
  name = property(lambda self: self._parameters['name'])
  chain_shard = property(lambda self: self._parameters['experimental_chain_shard'])

  @property
  def is_calibrated(self):
    return True

  def bootstrap_results(self, init_state):
    return self._impl.bootstrap_results(init_state)

  def _step_and_L(self, pkr):
    if self._parameters['store_parameters_in_results']:
      return pkr.accepted_results.step_size, int(pkr.accepted_results.num_leapfrog_steps)
    return self.step_size, int(self.num_leapfrog_steps)

  @property
  def _lockstep(self):
    return getattr(self._target, 'is_lockstep', False)

  @staticmethod
  def _momentum_scale(accepted_results, shapes, device):
    """sqrt of the diagonal inverse mass matrix ([D] CUDA tensor) when the results carry a momentum distribution
    (PreconditionedHamiltonianMonteCarlo), else None."""
    md = getattr(accepted_results, 'momentum_distribution', None)
    return None if md is None else md.scale_vector(shapes, device)

  def one_step(self, current_state, previous_kernel_results, seed=None):
    pkr = previous_kernel_results
    if self._lockstep:   # MetropolisHastings(UncalibratedHMC) over the lock-step leapfrog
      return self._lockstep_one_step(current_state, pkr, seed)
    seed = pb_random.sanitize_seed(seed)
    x, shapes, was_list = _engine.flatten_state(current_state)
    x = x.clone()
    B, D = x.shape
    acc = pkr.accepted_results
    g = _engine.flatten_state(list(acc.grads_target_log_prob))[0].clone()
    lp = acc.target_log_prob.contiguous().clone()
    step_size, L = self._step_and_L(pkr)
    step, step_kind = _engine.step_size_tensor(step_size, B, D, shapes, x.device)
    want = ('log_accept_ratio', 'is_accepted', 'proposed_state', 'proposed_target_log_prob',
            'proposed_grads', 'log_acceptance_correction', 'initial_momentum', 'final_momentum')
    out, _, _ = _engine.run(self._target, x, lp, g, step, step_kind, shapes, kind=_lib.KERNEL_HMC,
                            num_results=1, step_seeds=seed[None, :], num_leapfrog_steps=L, want=want,
                            shard=self.chain_shard, momentum_scale=self._momentum_scale(acc, shapes, x.device))
    is_acc = out['is_accepted'][0]
    proposal_seed = pb_random.split_seed(seed)[0]
    proposed = acc._replace(
        log_acceptance_correction=out['log_acceptance_correction'][0],
        target_log_prob=out['proposed_target_log_prob'][0],
        grads_target_log_prob=_engine.unflatten(out['proposed_grads'][0], shapes, True),
        initial_momentum=_engine.unflatten(out['initial_momentum'][0], shapes, was_list),
        final_momentum=_engine.unflatten(out['final_momentum'][0], shapes, was_list), seed=proposal_seed)
    accepted = choose(is_acc, _strip_seed(proposed), acc)
    results = MetropolisHastingsKernelResults(
        accepted_results=accepted, is_accepted=is_acc, log_accept_ratio=out['log_accept_ratio'][0],
        proposed_state=_engine.unflatten(out['proposed_state'][0], shapes, was_list),
        proposed_results=proposed, extra=[], seed=seed)
    return _engine.unflatten(x, shapes, was_list), results

  def _lockstep_one_step(self, current_state, pkr, seed):
    """Lock-step targets (row-sharded data, tensor-core logistic): the same transition as
    MetropolisHastings(UncalibratedHMC).one_step -- identical seeds and draws -- in three C-ABI calls: momentum draw,
    all L leapfrogs (the target's lock-step integrator), and ONE kernel for the Metropolis-Hastings step
    (pb2_hmc_mh_finish: correction, ratio, uniform, accept, choose of every field)."""
    import torch
    if type(self) is not HamiltonianMonteCarlo or self.chain_shard is not None:
      return self._impl.one_step(current_state, pkr, seed=seed)   # preconditioned / sharded variants: composed path
    seed = pb_random.sanitize_seed(seed)
    proposal_seed, acceptance_seed = pb_random.split_seed(seed)           # metropolis_hastings.py:183
    acc = pkr.accepted_results
    x, shapes, was_list = _engine.flatten_state(current_state)
    x = x.contiguous()
    B, D = x.shape
    g = _engine.flatten_state(list(acc.grads_target_log_prob))[0].contiguous()
    lp = acc.target_log_prob.contiguous()
    step_size, L = self._step_and_L(pkr)
    step, step_kind = _engine.step_size_tensor(step_size, B, D, shapes, x.device)
    sizes = _engine.part_sizes_of(shapes)
    seeds = pb_random.split_seed(proposal_seed, n=len(sizes))             # hmc.py:685
    m0 = torch.cat([pb_random.normal((B, n), seed=seeds[i], device=x.device) for i, n in enumerate(sizes)],
                   dim=1).contiguous()                                    # hmc.py:689-695
    m1, x1, lp1, g1 = self._target.leapfrog(m0, x, lp, g, step, step_kind, L)
    pm0 = _engine.flatten_state(acc.initial_momentum)[0].contiguous()
    pm1 = _engine.flatten_state(acc.final_momentum)[0].contiguous()
    pcorr = acc.log_acceptance_correction.contiguous()
    x_out, g_out, m0_out, m1_out = (torch.empty_like(x) for _ in range(4))
    lp_out, corr_acc, corr, ratio = (torch.empty_like(lp) for _ in range(4))
    is_acc = torch.empty(B, dtype=torch.uint8, device=x.device)
    ctx = _lib.Context.get(x.device)
    ctx.bind_stream()
    key = np.ascontiguousarray(acceptance_seed, np.uint32)
    _lib.check(ctx.lib.pb2_hmc_mh_finish(
        ctx.handle, B, D, B, 0, pb_random.default_layout(), _lib.u32p(key), _lib.ptr(m0), _lib.ptr(m1), _lib.ptr(x),
        _lib.ptr(lp), _lib.ptr(g), _lib.ptr(x1), _lib.ptr(lp1), _lib.ptr(g1), _lib.ptr(pm0), _lib.ptr(pm1),
        _lib.ptr(pcorr), _lib.ptr(x_out), _lib.ptr(lp_out), _lib.ptr(g_out), _lib.ptr(m0_out), _lib.ptr(m1_out),
        _lib.ptr(corr_acc), _lib.ptr(corr), _lib.ptr(ratio), _lib.ptr(is_acc)), ctx.handle)
    is_acc = is_acc.bool()
    proposed = acc._replace(
        log_acceptance_correction=corr, target_log_prob=lp1,
        grads_target_log_prob=_engine.unflatten(g1, shapes, True),
        initial_momentum=_engine.unflatten(m0, shapes, was_list),
        final_momentum=_engine.unflatten(m1, shapes, was_list), seed=proposal_seed)
    accepted = acc._replace(
        log_acceptance_correction=corr_acc, target_log_prob=lp_out,
        grads_target_log_prob=_engine.unflatten(g_out, shapes, True),
        initial_momentum=_engine.unflatten(m0_out, shapes, was_list),
        final_momentum=_engine.unflatten(m1_out, shapes, was_list), seed=[])
    results = MetropolisHastingsKernelResults(
        accepted_results=accepted, is_accepted=is_acc, log_accept_ratio=ratio,
        proposed_state=_engine.unflatten(x1, shapes, was_list), proposed_results=proposed, extra=[], seed=seed)
    return _engine.unflatten(x_out, shapes, was_list), results

  # ---- fused multi-transition driver used by sample_chain ------------------
  _FUSED_FIELDS = {
      ('is_accepted',): 'is_accepted', ('log_accept_ratio',): 'log_accept_ratio',
      ('accepted_results', 'target_log_prob'): 'target_log_prob',
      ('accepted_results', 'grads_target_log_prob'): 'grads_target_log_prob',
      ('accepted_results', 'step_size'): 'step_size',
      ('proposed_state',): 'proposed_state',
      ('proposed_results', 'target_log_prob'): 'proposed_target_log_prob',
      ('proposed_results', 'log_acceptance_correction'): 'log_acceptance_correction',
  }

  def _fused_run(self, x, shapes, was_list, pkr, seed, num_results, num_burnin_steps,
                 num_steps_between_results, paths, da_state=None, step=None, leapfrog_total=None,
                 da_over_ranks=False):
    """Runs all transitions in libpb2; returns (trace dict keyed by results path, final results, seed)."""
    if self._lockstep:
      return None     # row-sharded target: the step loop drives the per-leapfrog all-reduce
    acc = pkr.accepted_results
    B, D = x.shape
    g = _engine.flatten_state(list(acc.grads_target_log_prob))[0].clone()
    lp = acc.target_log_prob.contiguous().clone()
    step_size, L = self._step_and_L(pkr)
    if step is None:
      step, step_kind = _engine.step_size_tensor(step_size, B, D, shapes, x.device)
    else:
      step_kind = _lib.STEP_SCALAR
    want = {'states'}
    for p in paths:
      if p not in self._FUSED_FIELDS:
        return None
      want.add(self._FUSED_FIELDS[p])
    out, seed_out, step_seeds = _engine.run(
        self._target, x, lp, g, step, step_kind, shapes, kind=_lib.KERNEL_HMC, num_results=num_results,
        num_burnin_steps=num_burnin_steps, num_steps_between_results=num_steps_between_results, seed=seed,
        num_leapfrog_steps=L, want=tuple(want), da_state=da_state, shard=self.chain_shard,
        leapfrog_total=leapfrog_total, da_over_ranks=da_over_ranks,
        momentum_scale=self._momentum_scale(acc, shapes, x.device))
    traced = {}
    for p in paths:
      v = out[self._FUSED_FIELDS[p]] if self._FUSED_FIELDS[p] in out else None
      if p[-1] == 'grads_target_log_prob':
        v = _engine.unflatten(v, shapes, True)
      elif p[-1] == 'proposed_state':
        v = _engine.unflatten(v, shapes, was_list)
      traced[p] = v
    final_acc = acc._replace(target_log_prob=lp, grads_target_log_prob=_engine.unflatten(g, shapes, True))
    if self._parameters['store_parameters_in_results'] and step_kind == _lib.STEP_SCALAR:
      final_acc = final_acc._replace(step_size=step.reshape(()).clone())
    final = pkr._replace(accepted_results=final_acc, seed=step_seeds[-1].copy())
    return out['states'], traced, final, seed_out
