"""Sharded transition kernel (tfp/experimental/mcmc/sharded.py:22-90): independent chains on every member of a named
axis, obtained by folding the member's axis index into the step seed (distribute_lib.fold_in_axis_index, :193-207).

Named axes map onto `torch.distributed` process groups (probability_b200.distribute.register_axis); an unregistered
name means "the world": one member per rank, one process per GPU.  Chain sharding through `experimental_chain_shard`
is the other multi-GPU mode of this engine (one logical batch split over ranks, bit-identical to the unsharded run);
`Sharded` is the reference's: every rank owns its own batch and only the seed differs.
"""
from probability_b200 import distribute
from probability_b200 import random as pb_random
from probability_b200.mcmc import kernel as kernel_base


class Sharded(kernel_base.TransitionKernel):
  """Shards a transition kernel across a named axis."""

  def __init__(self, inner_kernel, chain_axis_names, validate_args=False, name=None):
    self._parameters = dict(
        inner_kernel=inner_kernel, chain_axis_names=distribute.canonicalize_named_axis(chain_axis_names),
        validate_args=validate_args, name=name)

  inner_kernel = property(lambda self: self._parameters['inner_kernel'])
  chain_axis_names = property(lambda self: self._parameters['chain_axis_names'])
  name = property(lambda self: self._parameters['name'])

  @property
  def is_calibrated(self):
    return self.inner_kernel.is_calibrated

  def bootstrap_results(self, init_state):
    return self.inner_kernel.bootstrap_results(init_state)

  def one_step(self, current_state, previous_kernel_results, seed=None):
    seed = pb_random.sanitize_seed(seed, salt='sharded_kernel')
    seed = distribute.fold_in_axis_index(seed, self.chain_axis_names)
    return self.inner_kernel.one_step(current_state, previous_kernel_results, seed=seed)

  @property
  def experimental_shard_axis_names(self):
    return self.inner_kernel.experimental_shard_axis_names

  def experimental_with_shard_axes(self, shard_axis_names):
    return self.copy(inner_kernel=self.inner_kernel.experimental_with_shard_axes(shard_axis_names))
