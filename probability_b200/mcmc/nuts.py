"""No-U-Turn Sampler (tfp/mcmc/nuts.py:165-1108) as one persistent CUDA transition
per chain: iterative tree doubling with popcount-indexed checkpoint stores, multinomial
sampling and the generalised U-turn criterion (MULTINOMIAL_SAMPLE = GENERALIZED_UTURN
= True, nuts.py:56-64)."""
import collections

import numpy as np

from probability_b200 import _lib
from probability_b200 import random as pb_random
from probability_b200.mcmc import _engine
from probability_b200.mcmc import kernel as kernel_base

NUTSKernelResults = collections.namedtuple(
    'NUTSKernelResults',
    ['target_log_prob', 'grads_target_log_prob', 'step_size', 'log_accept_ratio', 'leapfrogs_taken',
     'is_accepted', 'reach_max_depth', 'has_divergence', 'energy', 'seed'])


# experimental/mcmc/preconditioned_nuts.py:86-131: the same results plus the momentum distribution
PreconditionedNUTSKernelResults = collections.namedtuple(
    'PreconditionedNUTSKernelResults', NUTSKernelResults._fields + ('momentum_distribution',))


def _results_like(pkr, **fields):
  """NUTS results of the type of `pkr` (the preconditioned kernel carries its momentum distribution along)."""
  if hasattr(pkr, 'momentum_distribution'):
    return PreconditionedNUTSKernelResults(momentum_distribution=pkr.momentum_distribution, **fields)
  return NUTSKernelResults(**fields)


def momentum_scale_of(results, shapes, device):
  """sqrt of the diagonal inverse mass matrix as a [D] CUDA tensor, or None for the identity."""
  md = getattr(results, 'momentum_distribution', None)
  return None if md is None else md.scale_vector(shapes, device)


def build_tree_uturn_instruction(max_depth, init_memory=0):
  """(left, right) leaf index pairs of every balanced subtree (nuts.py:1013-1028)."""
  pairs = set()

  def rec(address, depth):
    if depth == 0:
      return address + 1, address + 1
    left, right = rec(address, depth - 1)
    _, right = rec(right, depth - 1)
    pairs.add((left, right))
    return left, right

  rec(init_memory, max_depth)
  return np.array(sorted(pairs), dtype=np.int32)


def generate_efficient_write_read_instruction(max_tree_depth):
  """Closed form of nuts.py:1031-1071 (what the CUDA kernel evaluates with popc/ffs):
  even step i writes slot popcount(i), odd steps write the trash slot `max_tree_depth`;
  odd step i reads slots [popcount(i) - trailing_ones(i), popcount(i))."""
  n = 1 << max_tree_depth
  write = np.zeros(n, np.int32)
  read = np.zeros((n, 2), np.int32)
  for i in range(n):
    pc = bin(i).count('1')
    if i % 2 == 0:
      write[i] = pc
    else:
      write[i] = max_tree_depth
      t = 0
      while (i >> t) & 1:
        t += 1
      read[i] = (pc - t, pc)
  return write, read


class NoUTurnSampler(kernel_base.TransitionKernel):

  def __init__(self, target_log_prob_fn, step_size, max_tree_depth=10, max_energy_diff=1000.,
               unrolled_leapfrog_steps=1, parallel_iterations=10, experimental_shard_axis_names=None,
               experimental_chain_shard=None, name=None):
    max_tree_depth = int(max_tree_depth)
    if max_tree_depth < 1:
      raise ValueError('max_tree_depth must be >= 1 but was {}'.format(max_tree_depth))
    if max_tree_depth > 12:
      raise ValueError('max_tree_depth > 12 is not supported by the CUDA kernel')
    self._parameters = dict(
        target_log_prob_fn=target_log_prob_fn, step_size=step_size, max_tree_depth=max_tree_depth,
        max_energy_diff=max_energy_diff, unrolled_leapfrog_steps=unrolled_leapfrog_steps,
        parallel_iterations=parallel_iterations, experimental_shard_axis_names=experimental_shard_axis_names,
        experimental_chain_shard=experimental_chain_shard, name=name)
    self._target = _engine.require_target(target_log_prob_fn)
    _engine.check_shard_axis_names(experimental_shard_axis_names)
    self._write_instruction, self._read_instruction = generate_efficient_write_read_instruction(
        max_tree_depth)

  target_log_prob_fn = property(lambda self: self._parameters['target_log_prob_fn'])
  step_size = property(lambda self: self._parameters['step_size'])
  max_tree_depth = property(lambda self: self._parameters['max_tree_depth'])
  max_energy_diff = property(lambda self: self._parameters['max_energy_diff'])
  unrolled_leapfrog_steps = property(lambda self: self._parameters['unrolled_leapfrog_steps'])
  name = property(lambda self: self._parameters['name'])
  chain_shard = property(lambda self: self._parameters['experimental_chain_shard'])
  write_instruction = property(lambda self: self._write_instruction)
  read_instruction = property(lambda self: self._read_instruction)

  @property
  def is_calibrated(self):
    return True

  def bootstrap_results(self, init_state):
    import torch
    if _engine.is_list_like(init_state):
      for p in init_state:
        if _engine.is_list_like(p) or isinstance(p, dict):
          raise TypeError('NUTS does not currently support nested or non-list-like state structures '
                          '(saw: {}).'.format(init_state))
    elif isinstance(init_state, dict):
      raise TypeError('NUTS does not currently support nested or non-list-like state structures '
                      '(saw: {}).'.format(init_state))
    x, shapes, _ = _engine.flatten_state(init_state)
    B, D = x.shape
    _engine.step_size_tensor(self.step_size, B, D, shapes, x.device)   # validates (nuts.py:472-473)
    lp, g = self._target.log_prob_and_grad(x)
    energy = lp - 0.5 * float(D)                                       # dummy momentum of ones (:461,498-500)
    return NUTSKernelResults(
        target_log_prob=lp, grads_target_log_prob=_engine.unflatten(g, shapes, True),
        step_size=_as_tensor_struct(self.step_size, x.device),
        log_accept_ratio=torch.zeros_like(lp),
        leapfrogs_taken=torch.zeros_like(lp, dtype=torch.int32),
        is_accepted=torch.zeros_like(lp, dtype=torch.bool),
        reach_max_depth=torch.zeros_like(lp, dtype=torch.bool),
        has_divergence=torch.zeros_like(lp, dtype=torch.bool), energy=energy,
        seed=pb_random.zeros_seed())

  _ALL = ('target_log_prob', 'grads_target_log_prob', 'log_accept_ratio', 'leapfrogs_taken', 'is_accepted',
          'reach_max_depth', 'has_divergence', 'energy')

  def one_step(self, current_state, previous_kernel_results, seed=None):
    pkr = previous_kernel_results
    seed = pb_random.sanitize_seed(seed)
    if _engine.is_list_like(current_state):
      for p in current_state:
        if _engine.is_list_like(p) or isinstance(p, dict):
          raise TypeError('NUTS does not currently support nested or non-list-like state structures '
                          '(saw: {}).'.format(current_state))
    x, shapes, was_list = _engine.flatten_state(current_state)
    x = x.clone()
    B, D = x.shape
    g = _engine.flatten_state(list(pkr.grads_target_log_prob))[0].clone()
    lp = pkr.target_log_prob.contiguous().clone()
    step, step_kind = _engine.step_size_tensor(pkr.step_size, B, D, shapes, x.device)
    out, _, _ = _engine.run(
        self._target, x, lp, g, step, step_kind, shapes, kind=_lib.KERNEL_NUTS, num_results=1,
        step_seeds=seed[None, :], max_tree_depth=self.max_tree_depth, max_energy_diff=self.max_energy_diff,
        unrolled_leapfrog_steps=self.unrolled_leapfrog_steps,
        want=('log_accept_ratio', 'leapfrogs_taken', 'is_accepted', 'reach_max_depth', 'has_divergence',
              'energy'), shard=self.chain_shard, momentum_scale=momentum_scale_of(pkr, shapes, x.device))
    results = _results_like(
        pkr,
        target_log_prob=lp, grads_target_log_prob=_engine.unflatten(g, shapes, True),
        step_size=pkr.step_size, log_accept_ratio=out['log_accept_ratio'][0],
        leapfrogs_taken=out['leapfrogs_taken'][0], is_accepted=out['is_accepted'][0],
        reach_max_depth=out['reach_max_depth'][0], has_divergence=out['has_divergence'][0],
        energy=out['energy'][0], seed=seed)
    return _engine.unflatten(x, shapes, was_list), results

  _FUSED_FIELDS = {(f,): f for f in _ALL + ('step_size',)}

  def _fused_run(self, x, shapes, was_list, pkr, seed, num_results, num_burnin_steps,
                 num_steps_between_results, paths, da_state=None, step=None, leapfrog_total=None,
                 da_over_ranks=False):
    B, D = x.shape
    g = _engine.flatten_state(list(pkr.grads_target_log_prob))[0].clone()
    lp = pkr.target_log_prob.contiguous().clone()
    if step is None:
      step, step_kind = _engine.step_size_tensor(pkr.step_size, B, D, shapes, x.device)
    else:
      step_kind = _lib.STEP_SCALAR
    want = {'states', 'log_accept_ratio', 'leapfrogs_taken', 'is_accepted', 'reach_max_depth',
            'has_divergence', 'energy'}
    for p in paths:
      if p not in self._FUSED_FIELDS:
        return None
      want.add(self._FUSED_FIELDS[p])
    out, seed_out, step_seeds = _engine.run(
        self._target, x, lp, g, step, step_kind, shapes, kind=_lib.KERNEL_NUTS, num_results=num_results,
        num_burnin_steps=num_burnin_steps, num_steps_between_results=num_steps_between_results, seed=seed,
        max_tree_depth=self.max_tree_depth, max_energy_diff=self.max_energy_diff,
        unrolled_leapfrog_steps=self.unrolled_leapfrog_steps, want=tuple(want), da_state=da_state,
        shard=self.chain_shard, leapfrog_total=leapfrog_total, da_over_ranks=da_over_ranks,
        momentum_scale=momentum_scale_of(pkr, shapes, x.device))
    traced = {}
    for p in paths:
      v = out.get(self._FUSED_FIELDS[p])
      if p[-1] == 'grads_target_log_prob':
        v = _engine.unflatten(v, shapes, True)
      traced[p] = v
    new_step = pkr.step_size
    if step_kind == _lib.STEP_SCALAR and da_state is not None:
      new_step = step.reshape(()).clone()
    final = _results_like(
        pkr,
        target_log_prob=lp, grads_target_log_prob=_engine.unflatten(g, shapes, True), step_size=new_step,
        log_accept_ratio=out['log_accept_ratio'][-1], leapfrogs_taken=out['leapfrogs_taken'][-1],
        is_accepted=out['is_accepted'][-1], reach_max_depth=out['reach_max_depth'][-1],
        has_divergence=out['has_divergence'][-1], energy=out['energy'][-1], seed=step_seeds[-1].copy())
    return out['states'], traced, final, seed_out


def _as_tensor_struct(step_size, device):
  import torch
  if _engine.is_list_like(step_size):
    return [torch.as_tensor(s, dtype=torch.float32, device=device) for s in step_size]
  return torch.as_tensor(step_size, dtype=torch.float32, device=device)
