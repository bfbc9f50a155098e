"""world_size-2 gloo test (CPU): the cross-rank pieces of the N>1 path that do not need a GPU --
(i) combining per-rank fixed-point partial sums of the accept probabilities reproduces the unsharded
reduce_logmeanexp that dual averaging uses (distribute_lib.py:147-162), exactly for any split, (ii) chain shards driven by the same seed with
global-chain-index counters reproduce the unsharded transition (oracle), (iii) row shards: the
all-reduced packed (gradient | log-lik) equals the full-data value."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  from oracle import mcmc as omcmc
  from oracle import rng as orng
  from oracle import targets as otargets
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    rng = np.random.default_rng(0)
    Bg = 16
    B = Bg // world
    # (i) partial log-mean-exp combination, as pb2_da_partial / pb2_da_apply do it
    lar = rng.standard_normal(Bg).astype(np.float32) - 0.5
    lar[3] = np.nan
    mine = lar[rank * B:(rank + 1) * B]
    la = np.minimum(np.where(np.isfinite(mine), mine, -np.inf), 0).astype(np.float32)
    # the partial: the accept probabilities summed as 64-bit fixed point in 2^-36 units (integer addition is
    # associative: any split of the chains over ranks gives the same bits), gathered as raw bytes
    fixed = lambda v: int(np.sum((np.exp(v).astype(np.float32) * np.float32(2.0 ** 36)).astype(np.uint64)))
    part = torch.tensor([fixed(la)], dtype=torch.int64)
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, part)
    total = int(torch.stack(gathered).sum())
    la_all = np.minimum(np.where(np.isfinite(lar), lar, -np.inf), 0).astype(np.float32)
    assert total == fixed(la_all)                       # sharded == unsharded, exactly
    combined = np.log(total / 2.0 ** 36 / Bg)
    expect = omcmc.DualAveraging.reduce_logmeanexp(la_all)
    np.testing.assert_allclose(combined, expect, rtol=1e-6)
    # (ii) chain shards == unsharded (NUTS + HMC), then all ranks agree on the gathered result
    es = otargets.EightSchools()
    x0 = np.tile(np.array([0, 0] + [1] * 8, np.float32), (Bg, 1)) + 0.1 * rng.standard_normal((Bg, 10)).astype(np.float32)
    lp0, g0 = es.logp_grad(x0)
    k = orng.key(5)
    full = omcmc.nuts_one_step(es, x0, lp0, g0, 0.3, k, max_tree_depth=5)
    sl = slice(rank * B, (rank + 1) * B)
    sh = omcmc.nuts_one_step(es, x0[sl], lp0[sl], g0[sl], 0.3, k, max_tree_depth=5, chain_offset=rank * B, B_global=Bg)
    np.testing.assert_array_equal(sh['state'], full['state'][sl])
    np.testing.assert_array_equal(sh['leapfrogs_taken'], full['leapfrogs_taken'][sl])
    # (iii) row shards: all-reduce(sum) of per-rank likelihood gradients == full data
    X, y = otargets.synthetic_logistic_data(400, 6, seed=1)
    th = (0.2 * rng.standard_normal((5, 7))).astype(np.float64)
    per = 400 // world
    lo, hi = rank * per, (rank + 1) * per
    def lik(Xs, ys):
      z = th @ Xs.T.astype(np.float64)
      w = ys[None, :] - 1 / (1 + np.exp(-z))
      return np.concatenate([w @ Xs, (ys[None, :] * z - np.logaddexp(0, z)).sum(1, keepdims=True)], 1)
    packed = torch.tensor(lik(X[lo:hi], y[lo:hi]))
    dist.all_reduce(packed)
    np.testing.assert_allclose(packed.numpy(), lik(X, y), rtol=1e-10)
    # (iv) Sharded / fold_in_axis_index: an unregistered axis name is the world -- every rank folds its own rank into
    # the key (distribute_lib.py:193-207), so the ranks' keys differ and match the oracle's fold_in
    from probability_b200 import distribute
    mine = distribute.fold_in_axis_index(orng.key(9), 'ranks')
    np.testing.assert_array_equal(mine, orng.fold_in(orng.key(9), rank))
    assert distribute.get_axis_size('ranks') == world
    ks = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(ks, torch.tensor(mine.astype(np.int64)))
    assert len({tuple(k.tolist()) for k in ks}) == world
    q.put((rank, 'ok'))
  except Exception as e:  # pylint: disable=broad-except
    q.put((rank, 'FAIL: %r' % (e,)))
  finally:
    dist.destroy_process_group()


def test_gloo_world2():
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = 29400 + (os.getpid() % 500)
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=300) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(res) == [(0, 'ok'), (1, 'ok')], res
