"""Named-axis helpers and the Sharded kernel's seed handling on the host (no GPU):
distribute_lib.canonicalize_named_axis / fold_in_axis_index (tfp/internal/distribute_lib.py:33-73,193-207) and
experimental/mcmc/sharded.py:63-74, checked against the oracle's threefry fold_in."""
import numpy as np
import pytest

from oracle import rng as orng


@pytest.fixture()
def dist_mod():
  from probability_b200 import distribute
  yield distribute
  for name in ('chains', 'data', 'model'):
    distribute.unregister_axis(name)


def test_canonicalize_named_axis(dist_mod):
  assert dist_mod.canonicalize_named_axis(None) == []
  assert dist_mod.canonicalize_named_axis('a') == ['a']
  assert dist_mod.canonicalize_named_axis(('a', 'b')) == ['a', 'b']


def test_fold_in_axis_index_matches_oracle_and_separates_members(dist_mod):
  from probability_b200 import random as pb_random
  seed = orng.key(1234)
  keys = []
  for idx in range(8):
    dist_mod.register_axis('chains', idx, 8)
    k = dist_mod.fold_in_axis_index(seed, 'chains')
    np.testing.assert_array_equal(k, orng.fold_in(seed, idx))
    keys.append(tuple(int(v) for v in k))
  assert len(set(keys)) == 8
  # several axes fold in order; None leaves the seed alone; an unregistered axis without a process group is index 0
  dist_mod.register_axis('chains', 3, 8)
  dist_mod.register_axis('data', 1, 2)
  np.testing.assert_array_equal(dist_mod.fold_in_axis_index(seed, ['chains', 'data']),
                                orng.fold_in(orng.fold_in(seed, 3), 1))
  np.testing.assert_array_equal(dist_mod.fold_in_axis_index(seed, None), seed)
  np.testing.assert_array_equal(dist_mod.fold_in_axis_index(seed, 'nowhere'), orng.fold_in(seed, 0))
  assert dist_mod.get_axis_size('data') == 2 and dist_mod.get_axis_index('data') == 1
  del pb_random


def test_sharded_kernel_folds_salt_then_axis_index(dist_mod):
  """Sharded.one_step: sanitize_seed(seed, salt='sharded_kernel') then fold_in_axis_index (sharded.py:63-67); the
  inner kernel here just records the seed it is handed."""
  from probability_b200 import mcmc
  from probability_b200 import random as pb_random

  class Recorder(mcmc.TransitionKernel):
    is_calibrated = False

    def bootstrap_results(self, init_state):
      return ()

    def one_step(self, current_state, previous_kernel_results, seed=None):
      return np.asarray(seed), previous_kernel_results

  got = []
  for idx in range(2):
    dist_mod.register_axis('chains', idx, 2)
    k = mcmc.Sharded(Recorder(), 'chains')
    assert k.chain_axis_names == ['chains'] and k.is_calibrated is False
    s, _ = k.one_step(0.0, k.bootstrap_results(0.0), seed=orng.key(7))
    np.testing.assert_array_equal(s, orng.fold_in(pb_random.sanitize_seed(orng.key(7), salt='sharded_kernel'), idx))
    got.append(tuple(int(v) for v in s))
  assert got[0] != got[1]
  with pytest.raises(ValueError):
    mcmc.Sharded(Recorder(), 'chains').one_step(0.0, (), seed=None)


def test_state_part_shard_axes_are_size_one_only(dist_mod):
  from probability_b200.mcmc import _engine
  dist_mod.register_axis('model', 0, 1)
  _engine.check_shard_axis_names(['model', None])
  _engine.check_shard_axis_names(None)
  dist_mod.register_axis('model', 0, 4)
  with pytest.raises(NotImplementedError):
    _engine.check_shard_axis_names([['model'], None])


def test_merged_running_moments_equal_the_pooled_moments():
  """experimental.stats.merge_moment_states: Chan's merge of per-rank [count, mean, M2] states == the moments of the
  pooled observations (what windowed adaptation uses when the chains are sharded over ranks)."""
  import torch
  from probability_b200.experimental.stats import merge_moment_states
  rng = np.random.default_rng(3)
  D = 7
  chunks = [rng.standard_normal((n, D)) * np.arange(1, D + 1) + 3.0 for n in (5, 1, 40, 17)]
  states = []
  for c in chunks:
    m = c.mean(0)
    states.append(torch.tensor(np.concatenate([[len(c)], m, ((c - m) ** 2).sum(0)]), dtype=torch.float32))
  states.insert(2, torch.zeros(1 + 2 * D))              # a rank without observations
  got = merge_moment_states(states).numpy()
  pooled = np.concatenate(chunks)
  assert got[0] == len(pooled)
  np.testing.assert_allclose(got[1:1 + D], pooled.mean(0), rtol=1e-5)
  np.testing.assert_allclose(got[1 + D:] / got[0], pooled.var(0), rtol=1e-5)
  np.testing.assert_array_equal(merge_moment_states([states[0]]).numpy(), states[0].numpy())
