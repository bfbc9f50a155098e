"""TransitionKernel base class (tfp/mcmc/kernel.py:27-164)."""
import abc


class TransitionKernel(metaclass=abc.ABCMeta):
  """`one_step(current_state, previous_kernel_results, seed=None)` ->
  `(next_state, kernel_results)`; `bootstrap_results(init_state)`; `is_calibrated`."""

  @abc.abstractmethod
  def one_step(self, current_state, previous_kernel_results, seed=None):
    """Takes one step of the TransitionKernel."""

  @abc.abstractproperty
  def is_calibrated(self):
    """Returns `True` if Markov chain converges to specified distribution."""

  def bootstrap_results(self, init_state):
    raise NotImplementedError()

  @property
  def parameters(self):
    return getattr(self, '_parameters', {})

  def copy(self, **override_parameter_kwargs):
    """Non-destructively creates a deep copy of the kernel (kernel.py:148-164)."""
    parameters = dict(self.parameters, **override_parameter_kwargs)
    new_kernel = type(self)(**parameters)
    return new_kernel

  @property
  def experimental_shard_axis_names(self):
    """Named axes over which the state parts are sharded (kernel.py:117-126); none unless the kernel takes the
    `experimental_shard_axis_names` parameter."""
    return self.parameters.get('experimental_shard_axis_names') or []

  def experimental_with_shard_axes(self, shard_axis_names):
    """A copy of the kernel whose state parts are sharded over `shard_axis_names` (hmc.py:686-687, nuts.py:516-517)."""
    if 'experimental_shard_axis_names' in self.parameters:
      return self.copy(experimental_shard_axis_names=shard_axis_names)
    return self
