"""Sharded kernel on the GPU (tfp/experimental/mcmc/sharded.py:22-90; reference test
sharded_test.py:45-59 `test_sharded_kernel_produces_independent_chains`): every member of the named axis runs the
inner kernel with its axis index folded into the seed, so the members' chains differ, and each member reproduces the
oracle driven by that folded seed."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu

from oracle import mcmc as omcmc  # noqa: E402
from oracle import rng as orng  # noqa: E402
from oracle import targets as otargets  # noqa: E402


@pytest.fixture(scope='module')
def tfp():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  import probability_b200 as tfp_
  return tfp_


def test_sharded_kernel_produces_independent_chains_and_matches_oracle(tfp):
  dev = torch.device('cuda', 0)
  tg = tfp.targets.EightSchools()
  es = otargets.EightSchools()
  rng = np.random.default_rng(0)
  x = (np.array([0, 0] + [1] * 8) + 0.3 * rng.standard_normal((64, 10))).astype(np.float32)
  xt = torch.tensor(x, device=dev)
  state = [xt[:, 0].contiguous(), xt[:, 1].contiguous(), xt[:, 2:].contiguous()]
  lp0, g0 = es.logp_grad(x)
  seed = orng.key(3)
  outs = []
  try:
    for idx in range(3):
      tfp.distribute.register_axis('chains', idx, 3)
      k = tfp.mcmc.Sharded(tfp.mcmc.NoUTurnSampler(tg, step_size=0.3, max_tree_depth=5), 'chains')
      new_state, kr = k.one_step(state, k.bootstrap_results(state), seed=seed)
      got = torch.cat([p.reshape(64, -1) for p in new_state], 1).cpu().numpy()
      folded = orng.fold_in(tfp.random.sanitize_seed(seed, salt='sharded_kernel'), idx)
      ref = omcmc.nuts_one_step(es, x, lp0, g0, 0.3, folded, max_tree_depth=5)
      same = kr.leapfrogs_taken.cpu().numpy() == ref['leapfrogs_taken']
      assert same.mean() >= 0.97
      np.testing.assert_allclose(got[same], ref['state'][same], rtol=2e-3, atol=2e-3)
      outs.append(got)
  finally:
    tfp.distribute.unregister_axis('chains')
  for i in range(3):
    for j in range(i + 1, 3):
      assert not np.allclose(outs[i], outs[j])


def test_with_shard_axes_round_trip(tfp):
  tg = tfp.targets.EightSchools()
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.3)
  assert k.experimental_shard_axis_names == []
  tfp.distribute.register_axis('model', 0, 1)
  try:
    k2 = tfp.mcmc.DualAveragingStepSizeAdaptation(k, num_adaptation_steps=3).experimental_with_shard_axes(['model', None, None])
    assert k2.experimental_shard_axis_names == ['model', None, None]
    assert k2.inner_kernel.experimental_shard_axis_names == ['model', None, None]
    s = tfp.mcmc.Sharded(k, 'chains').experimental_with_shard_axes(['model', None, None])
    assert s.experimental_shard_axis_names == ['model', None, None]
  finally:
    tfp.distribute.unregister_axis('model')
