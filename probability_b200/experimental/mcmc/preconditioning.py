"""Diagonally preconditioned HMC / NUTS (tfp/experimental/mcmc/preconditioned_hmc.py:42-330,
preconditioned_nuts.py:169-330, preconditioning_utils.py).

The reference takes a `momentum_distribution` (a tfd distribution whose covariance is the mass matrix) and evaluates
the kinetic energy / velocity through it.  This engine has no distribution library: the supported momentum
distributions are the diagonal Gaussians that DiagonalMassMatrixAdaptation builds, represented by `DiagonalMomentum`
(what `preconditioning_utils.make_momentum_distribution(state_parts, batch_shape, running_variance_parts)` returns
there: MultivariateNormalPrecisionFactorLinearOperator with precision factor diag(sqrt(variance))).  The CUDA kernels
realise the mass matrix as the change of variables u = x / sqrt(variance) (pb2_targets.cuh ScaledT; the dense Gaussian
folds it into its precision matrix and stays on the tensor cores); states, gradients and momenta cross the C ABI in
the original coordinates (pb2_run_cfg.d_momentum_scale).
"""
import collections

import numpy as np

from probability_b200.mcmc import _engine
from probability_b200.mcmc import hmc as hmc_lib
from probability_b200.mcmc import nuts as nuts_lib


class DiagonalMomentum(object):
  """Momentum ~ N(0, diag(1 / variance)) per state part: `variance` = diagonal of the INVERSE mass matrix (the running
  variance of the state), one tensor of the part's event shape per state part."""

  def __init__(self, variance_parts):
    import torch
    parts = list(variance_parts) if _engine.is_list_like(variance_parts) else [variance_parts]
    self.variance_parts = [torch.as_tensor(v, dtype=torch.float32) for v in parts]
    self._cache = {}

  def variance(self):
    return self.variance_parts

  def scale_vector(self, shapes, device):
    """sqrt(variance) flattened to the kernels' [D] layout."""
    import torch
    key = (tuple(shapes), str(device))
    if key not in self._cache:
      sizes = _engine.part_sizes_of(shapes)
      parts = self.variance_parts
      if len(parts) == 1 and len(sizes) > 1 and parts[0].numel() == sum(sizes):
        flat = parts[0].reshape(-1)
      else:
        if len(parts) != len(sizes):
          raise ValueError('the momentum distribution has {} parts but the state has {}'.format(len(parts), len(sizes)))
        cols = []
        for v, n, shp in zip(parts, sizes, shapes):
          if v.dim() > len(shp):
            raise NotImplementedError('a batch of mass matrices (one per chain) is not supported: the variance of a '
                                      'state part must have the part\'s event shape')
          cols.append(torch.broadcast_to(v.to(device), shp if len(shp) else (1,)).reshape(-1))
        flat = torch.cat(cols)
      self._cache[key] = torch.sqrt(flat.to(device=device, dtype=torch.float32)).contiguous()
    return self._cache[key]


def make_momentum_distribution(state_parts, batch_shape, running_variance_parts, shard_axis_names=None):
  """preconditioning_utils.make_momentum_distribution: the momentum distribution whose covariance is the inverse of the
  given (running) variances."""
  del state_parts, batch_shape, shard_axis_names
  return DiagonalMomentum(running_variance_parts)


def update_momentum_distribution(momentum_distribution, running_variance_parts):
  """preconditioning_utils.update_momentum_distribution."""
  del momentum_distribution
  return DiagonalMomentum(running_variance_parts)


UncalibratedPreconditionedHamiltonianMonteCarloKernelResults = collections.namedtuple(
    'UncalibratedPreconditionedHamiltonianMonteCarloKernelResults',
    hmc_lib.UncalibratedHamiltonianMonteCarloKernelResults._fields + ('momentum_distribution',))
PreconditionedNUTSKernelResults = nuts_lib.PreconditionedNUTSKernelResults


def _identity_momentum(init_state):
  import torch
  x, shapes, _ = _engine.flatten_state(init_state)
  return DiagonalMomentum([torch.ones(s if len(s) else (), device=x.device) for s in shapes])


class PreconditionedHamiltonianMonteCarlo(hmc_lib.HamiltonianMonteCarlo):
  """HamiltonianMonteCarlo whose momentum is drawn from `momentum_distribution` (preconditioned_hmc.py:42); None is
  the identity mass matrix.  The fused CUDA transition is HamiltonianMonteCarlo's; the momentum distribution travels in
  `accepted_results.momentum_distribution` so that DiagonalMassMatrixAdaptation can replace it."""

  def __init__(self, target_log_prob_fn, step_size, num_leapfrog_steps, momentum_distribution=None,
               state_gradients_are_stopped=False, store_parameters_in_results=False,
               experimental_shard_axis_names=None, experimental_chain_shard=None, name=None):
    super().__init__(target_log_prob_fn, step_size, num_leapfrog_steps,
                     state_gradients_are_stopped=state_gradients_are_stopped,
                     store_parameters_in_results=store_parameters_in_results,
                     experimental_shard_axis_names=experimental_shard_axis_names,
                     experimental_chain_shard=experimental_chain_shard, name=name)
    self._parameters['momentum_distribution'] = momentum_distribution
    if getattr(self._target, 'is_lockstep', False):
      raise NotImplementedError('row-sharded targets run the unpreconditioned lock-step leapfrog')

  momentum_distribution = property(lambda self: self._parameters['momentum_distribution'])

  def bootstrap_results(self, init_state):
    r = super().bootstrap_results(init_state)
    md = self.momentum_distribution or _identity_momentum(init_state)
    ext = lambda u: UncalibratedPreconditionedHamiltonianMonteCarloKernelResults(momentum_distribution=md,
                                                                                 **u._asdict())
    return r._replace(accepted_results=ext(r.accepted_results), proposed_results=ext(r.proposed_results))


class PreconditionedNoUTurnSampler(nuts_lib.NoUTurnSampler):
  """NoUTurnSampler with a diagonal mass matrix (preconditioned_nuts.py:169): momentum from `momentum_distribution`,
  positions move along the velocity, the U-turn criterion dots the cumulative momentum with the end velocities."""

  def __init__(self, target_log_prob_fn, step_size, max_tree_depth=10, max_energy_diff=1000.,
               unrolled_leapfrog_steps=1, parallel_iterations=10, momentum_distribution=None,
               experimental_shard_axis_names=None, experimental_chain_shard=None, name=None):
    super().__init__(target_log_prob_fn, step_size, max_tree_depth=max_tree_depth, max_energy_diff=max_energy_diff,
                     unrolled_leapfrog_steps=unrolled_leapfrog_steps, parallel_iterations=parallel_iterations,
                     experimental_shard_axis_names=experimental_shard_axis_names,
                     experimental_chain_shard=experimental_chain_shard, name=name)
    self._parameters['momentum_distribution'] = momentum_distribution

  momentum_distribution = property(lambda self: self._parameters['momentum_distribution'])

  def bootstrap_results(self, init_state):
    r = super().bootstrap_results(init_state)
    md = self.momentum_distribution or _identity_momentum(init_state)
    return PreconditionedNUTSKernelResults(momentum_distribution=md, **r._asdict())
