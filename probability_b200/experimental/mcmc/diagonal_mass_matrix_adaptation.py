"""DiagonalMassMatrixAdaptation (tfp/experimental/mcmc/diagonal_mass_matrix_adaptation.py:73-330): keeps a running
variance of the state (all chains of a step are observations; pb2_running_moments_update on the device) and sets the
inner kernel's momentum distribution to the diagonal Gaussian whose covariance is the INVERSE of it -- at every step
(`num_estimation_steps=None`) or once, when `step == num_estimation_steps`.  Not a calibrated sampler: one stage of
an adaptation schedule (windowed_sampling.py)."""
import collections

from probability_b200.experimental.mcmc import preconditioning
from probability_b200.experimental.stats import RunningVariance
from probability_b200.mcmc import _engine
from probability_b200.mcmc import kernel as kernel_base

DiagonalMassMatrixAdaptationResults = collections.namedtuple(
    'DiagonalMassMatrixAdaptationResults', ['inner_results', 'running_variance', 'num_estimation_steps', 'step'])


def hmc_like_momentum_distribution_setter_fn(kernel_results, new_distribution):
  """diagonal_mass_matrix_adaptation.py:37-52: the innermost results (under `accepted_results` for HMC)."""
  def rec(kr):
    if hasattr(kr, 'inner_results'):
      return kr._replace(inner_results=rec(kr.inner_results))
    if hasattr(kr, 'accepted_results'):
      out = kr._replace(accepted_results=kr.accepted_results._replace(momentum_distribution=new_distribution))
      if hasattr(kr, 'proposed_results') and hasattr(kr.proposed_results, 'momentum_distribution'):
        out = out._replace(proposed_results=out.proposed_results._replace(momentum_distribution=new_distribution))
      return out
    return kr._replace(momentum_distribution=new_distribution)
  return rec(kernel_results)


def hmc_like_momentum_distribution_getter_fn(kernel_results):
  kr = kernel_results
  while hasattr(kr, 'inner_results'):
    kr = kr.inner_results
  if hasattr(kr, 'accepted_results'):
    kr = kr.accepted_results
  return kr.momentum_distribution


class DiagonalMassMatrixAdaptation(kernel_base.TransitionKernel):

  def __init__(self, inner_kernel, initial_running_variance, num_estimation_steps=None,
               momentum_distribution_setter_fn=hmc_like_momentum_distribution_setter_fn,
               momentum_distribution_getter_fn=hmc_like_momentum_distribution_getter_fn, validate_args=False,
               experimental_shard_axis_names=None, name=None):
    self._parameters = dict(
        inner_kernel=inner_kernel, initial_running_variance=initial_running_variance,
        num_estimation_steps=num_estimation_steps, momentum_distribution_setter_fn=momentum_distribution_setter_fn,
        momentum_distribution_getter_fn=momentum_distribution_getter_fn, validate_args=validate_args,
        experimental_shard_axis_names=experimental_shard_axis_names, name=name)

  inner_kernel = property(lambda self: self._parameters['inner_kernel'])
  initial_running_variance = property(lambda self: self._parameters['initial_running_variance'])
  num_estimation_steps = property(lambda self: self._parameters['num_estimation_steps'])
  name = property(lambda self: self._parameters['name'])

  @property
  def is_calibrated(self):
    return False

  def momentum_distribution_setter_fn(self, kernel_results, new_momentum_distribution):
    return self._parameters['momentum_distribution_setter_fn'](kernel_results, new_momentum_distribution)

  def momentum_distribution_getter_fn(self, kernel_results):
    return self._parameters['momentum_distribution_getter_fn'](kernel_results)

  def _running_variance(self):
    rv = self.initial_running_variance
    if isinstance(rv, RunningVariance):
      return rv
    raise TypeError('initial_running_variance must be a probability_b200.experimental.stats.RunningVariance '
                    '(RunningVariance.from_shape / from_stats)')

  def bootstrap_results(self, init_state):
    inner_results = self.inner_kernel.bootstrap_results(init_state)
    results = self._bootstrap_from_inner_results(init_state, inner_results)
    if self.num_estimation_steps is not None:
      return results                     # the momentum is only updated at the end of the estimation phase (:296-300)
    md = preconditioning.DiagonalMomentum(results.running_variance.variance())
    return results._replace(inner_results=self.momentum_distribution_setter_fn(results.inner_results, md))

  def _bootstrap_from_inner_results(self, init_state, inner_results):
    del init_state
    n = -1 if self.num_estimation_steps is None else int(self.num_estimation_steps)
    return DiagonalMassMatrixAdaptationResults(inner_results=inner_results, running_variance=self._running_variance(),
                                               num_estimation_steps=n, step=0)

  def one_step(self, current_state, previous_kernel_results, seed=None):
    pkr = previous_kernel_results
    new_state, new_inner = self.inner_kernel.one_step(current_state, pkr.inner_results, seed=seed)
    step = int(pkr.step) + 1
    n_est = int(pkr.num_estimation_steps)
    every = self.num_estimation_steps is None
    rv = pkr.running_variance
    if every or step <= n_est:           # :280-284
      parts = list(new_state) if _engine.is_list_like(new_state) else [new_state]
      rv = rv.update(parts if rv.was_list else parts[0])
    if every or step == n_est:           # :285-287
      md = preconditioning.update_momentum_distribution(self.momentum_distribution_getter_fn(new_inner), rv.variance())
      new_inner = self.momentum_distribution_setter_fn(new_inner, md)
    return new_state, pkr._replace(inner_results=new_inner, running_variance=rv, step=step)
