"""C5: large-data Bayesian logistic regression (1e6 rows x 100 weights), rows sharded over the ranks, 1,024 replicated
chains, HMC with the per-leapfrog gradient all-reduced over NVLink (NCCL).  One JSON line (rank 0).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_c5.py
  (N = 1 works without torchrun)
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probability_b200 as tfp  # noqa: E402

N_ROWS, D, B, L, STEPS = 1_000_000, 100, 1024, 10, 10


def main():
  world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)
  per = N_ROWS // world
  lo = rank * per
  rng = np.random.default_rng(1000 + rank)          # each rank generates its own rows (same model everywhere)
  theta_true = np.random.default_rng(1).standard_normal(D).astype(np.float32) * 0.1
  X = rng.standard_normal((per, D - 1)).astype(np.float32)
  z = X @ theta_true[:-1] + theta_true[-1]
  y = (rng.random(per) < 1 / (1 + np.exp(-z))).astype(np.float32)
  out = {}
  for name, use_tc in (('tcgen05', True), ('fp32', False)):
    tg = tfp.targets.RowShardedLogisticRegression(X, y)
    tg.use_tensor_cores = use_tc
    k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=1.5e-3, num_leapfrog_steps=L)
    st = torch.zeros(B, D, device=dev)
    kr = k.bootstrap_results(st)
    for i in range(3):
      st, kr = k.one_step(st, kr, seed=(7, i))
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc = 0.0
    for i in range(STEPS):
      st, kr = k.one_step(st, kr, seed=(8, i))
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / STEPS
    out[name] = {'ms_per_transition': ms, 'chain_grad_evals_per_sec': B * L / ms * 1e3,
                 'tflops_algorithmic_all_gpus': 4.0 * N_ROWS * D * B * L / ms / 1e9,
                 'accept_rate': float(kr.is_accepted.float().mean())}
  if rank == 0:
    print(json.dumps({'config': 'C5 logistic regression %d rows x %d weights, %d chains, HMC L=%d, rows sharded over %d GPU(s), '
                                'gradient all-reduced every leapfrog' % (N_ROWS, D, B, L, world), 'n_gpus': world, **out}))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
