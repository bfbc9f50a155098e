"""Run under torchrun on >= 2 GPUs (one process per GPU, NCCL), once with the collectives issued from Python
(torch.distributed) and twice with the library-owned communicator (tfp.distribute.init_comm: fused sharded dual
averaging inside pb2_run; the per-leapfrog gradient sum inside pb2_rowshard_leapfrog once as an NCCL all-reduce between
the kernels and once as the peer-memory reduction fused into the step kernel):
  (a) chain-sharded NUTS + DualAveraging == the unsharded run (global-chain-index RNG counters, cross-rank
      log-mean-exp): same step sizes on every rank, same states bit for bit;
  (b) row-sharded logistic HMC with the per-leapfrog gradient all-reduce == all rows on one GPU, and all
      ranks (replicated chains) take identical accept decisions.
Prints 'MULTIGPU OK' on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import probability_b200 as tfp  # noqa: E402


def main():
  rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
  local = int(os.environ.get('LOCAL_RANK', rank))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  dist.init_process_group('nccl', device_id=dev)
  ctx = tfp._lib.Context.get(dev)
  for mode in ('torch.distributed collectives', 'library-owned communicator, NCCL all-reduce between the kernels',
               'library-owned communicator, peer-memory reduction inside the step kernel'):
    if mode.startswith('library'):
      assert tfp.distribute.init_comm() == world
      assert tfp.distribute.comm_size() == world
      t = torch.full((5,), float(rank + 1), device=dev)
      tfp.distribute.all_reduce_sum(t)
      assert torch.equal(t, torch.full((5,), world * (world + 1) / 2.0, device=dev))
      ctx.set_int('rowshard_collective', 1 if 'peer-memory' in mode else 0)
    check(rank, world, dev, mode)
  dist.barrier()
  if rank == 0:
    print('MULTIGPU OK world=%d' % world, flush=True)
  tfp.distribute.destroy_comm()
  dist.destroy_process_group()


def check(rank, world, dev, mode):
  # ---------------- (a) chain sharding
  Bg = 64 * world
  B = Bg // world
  tg = tfp.targets.EightSchools()
  rng = np.random.default_rng(0)
  x_all = (np.array([0, 0] + [1] * 8) + 0.3 * rng.standard_normal((Bg, 10))).astype(np.float32)

  def parts(x):
    x = torch.tensor(x, device=dev)
    return [x[:, 0].contiguous(), x[:, 1].contiguous(), x[:, 2:].contiguous()]

  def run(state, shard, axis):
    k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.1, max_tree_depth=5, experimental_chain_shard=shard)
    k = tfp.mcmc.DualAveragingStepSizeAdaptation(k, num_adaptation_steps=6,
                                                 experimental_reduce_chain_axis_names=axis)
    return tfp.mcmc.sample_chain(8, state, kernel=k, seed=11,
                                 trace_fn=lambda _, kr: (kr.inner_results.step_size, kr.inner_results.leapfrogs_taken))
  sharded = run(parts(x_all[rank * B:(rank + 1) * B]), tfp.mcmc.ChainShard(rank * B, Bg), 'ranks')
  full = run(parts(x_all), None, None)         # every rank also runs the whole job locally
  steps_sh = sharded.trace[0].cpu().numpy(); steps_full = full.trace[0].cpu().numpy()
  np.testing.assert_array_equal(steps_sh, steps_full)   # exact fixed-point accept statistic: bit-identical adaptation
  gathered = [torch.zeros_like(sharded.trace[0]) for _ in range(world)]
  dist.all_gather(gathered, sharded.trace[0].contiguous())
  for g in gathered:                           # same step size on every rank
    np.testing.assert_array_equal(g.cpu().numpy(), steps_sh)
  np.testing.assert_array_equal(sharded.trace[1].cpu().numpy(),
                                full.trace[1][:, rank * B:(rank + 1) * B].cpu().numpy())
  for a, b in zip(sharded.all_states, full.all_states):
    np.testing.assert_array_equal(a.cpu().numpy(), b[:, rank * B:(rank + 1) * B].cpu().numpy())

  # ---------------- (b) row sharding with per-leapfrog gradient all-reduce
  # (96 chains: FP32 thread-per-chain gradient; 256 chains: the tcgen05 gradient)
  n, d = 4000, 39
  for Bc in (96, 256):
    X, y = tfp.targets.synthetic_logistic_data(n, d, seed=1)
    per = n // world
    lo, hi = rank * per, (n if rank == world - 1 else (rank + 1) * per)
    t_shard = tfp.targets.RowShardedLogisticRegression(X[lo:hi], y[lo:hi])
    t_full = tfp.targets.RowShardedLogisticRegression(X, y)
    t_full._world = lambda: None                # all rows local: no collective
    st = torch.tensor((0.1 * np.random.default_rng(2).standard_normal((Bc, d + 1))).astype(np.float32), device=dev)
    outs = []
    for t in (t_shard, t_full):
      k = tfp.mcmc.HamiltonianMonteCarlo(t, step_size=0.01, num_leapfrog_steps=4)
      outs.append(tfp.mcmc.sample_chain(4, st, kernel=k, seed=5,
                                        trace_fn=lambda _, kr: (kr.is_accepted, kr.log_accept_ratio)))
    sh, fu = outs
    np.testing.assert_allclose(sh.trace[1].cpu().numpy(), fu.trace[1].cpu().numpy(), rtol=5e-3, atol=5e-3)
    agree = (sh.trace[0] == fu.trace[0]).float().mean().item()
    assert agree > 0.97, agree
    acc = sh.trace[0].to(torch.int32).contiguous()
    g = [torch.zeros_like(acc) for _ in range(world)]
    dist.all_gather(g, acc)
    for a in g:                                  # replicas take bit-identical decisions
      assert torch.equal(a, acc)
    s = [torch.zeros_like(sh.all_states) for _ in range(world)]
    dist.all_gather(s, sh.all_states.contiguous())
    for a in s:
      assert torch.equal(a, sh.all_states)
  # ---------------- (c) windowed adaptation with sharded chains: step size and mass matrix are adapted on ALL ranks'
  # chains (cross-rank accept statistic + rank-ordered merge of the running moments), so every rank ends with the same
  # step size and the same variance estimate, and those agree statistically with the unsharded run of the same batch
  tgw = tfp.targets.EightSchools()
  Bw = 64
  xw = (np.array([0, 0] + [1] * 8) + 0.3 * np.random.default_rng(5).standard_normal((Bw * world, 10))).astype(np.float32)
  import probability_b200.experimental.mcmc as emcmc
  dw, tw = emcmc.windowed_adaptive_nuts(20, tgw, n_chains=Bw, num_adaptation_steps=120, max_tree_depth=6,
                                        current_state=torch.tensor(xw[rank * Bw:(rank + 1) * Bw], device=dev),
                                        experimental_chain_shard=tfp.mcmc.ChainShard(rank * Bw, Bw * world), seed=9)
  flatv = lambda t: torch.cat([v.reshape(-1) for v in (t['variance_scaling'] if isinstance(t['variance_scaling'], (list, tuple))
                                                       else [t['variance_scaling']])])
  mine = torch.cat([tw['step_size'][:1].reshape(-1), flatv(tw)]).contiguous()
  gw = [torch.zeros_like(mine) for _ in range(world)]
  dist.all_gather(gw, mine)
  for a in gw:
    assert torch.equal(a, mine), 'ranks adapted to different step sizes / mass matrices'
  df, tf_ = emcmc.windowed_adaptive_nuts(20, tgw, n_chains=Bw * world, num_adaptation_steps=120, max_tree_depth=6,
                                         current_state=torch.tensor(xw, device=dev), seed=9)
  vs, vf = flatv(tw).cpu().numpy(), flatv(tf_).cpu().numpy()
  np.testing.assert_allclose(vs, vf, rtol=0.35)
  np.testing.assert_allclose(float(tw['step_size'][0]), float(tf_['step_size'][0]), rtol=0.35)
  dist.barrier()
  if rank == 0:
    print('multi-GPU invariants hold with %s' % mode, flush=True)


if __name__ == '__main__':
  main()
