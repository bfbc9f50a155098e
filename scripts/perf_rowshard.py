"""C5 (large-data logistic regression, rows sharded over 8 GPUs) on ONE rank's shard: time of the local gradient pass
for all chains and of a full HMC transition (without the all-reduce: one rank)."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
dev = torch.device('cuda', 0)
N, D, B = 125_000, 100, 1024      # rows per GPU at 8 GPUs, weights incl. bias, chains
rng = np.random.default_rng(1)
X = rng.standard_normal((N, D - 1)).astype(np.float32)
theta_true = rng.standard_normal(D).astype(np.float32) * 0.1
z = np.concatenate([X, np.ones((N, 1), np.float32)], 1) @ theta_true
y = (rng.random(N) < 1 / (1 + np.exp(-z))).astype(np.float32)
tg = tfp.targets.RowShardedLogisticRegression(X, y)
ctx = _lib.Context.get(dev); ctx.bind_stream()
Xd, yd = tg._device_data(dev)
th = torch.zeros(B, D, device=dev) + 0.01 * torch.randn(B, D, device=dev)
packed = torch.empty(B, D + 1, device=dev)
import ctypes
planes = tg._tc_planes(ctx, dev)
fp32 = lambda: ctx.lib.pb2_rowshard_logistic_grad(ctx.handle, _lib.ptr(Xd), _lib.ptr(yd), N, D, tg.padded_dim, _lib.ptr(th), B, _lib.ptr(packed))
tc = lambda: ctx.lib.pb2_rowshard_logistic_grad_tc(ctx.handle, _lib.ptr(planes), _lib.ptr(yd), N, D, _lib.ptr(th), B, _lib.ptr(packed))
for name, fn in (('fp32 thread-per-chain', fp32), ('tcgen05', tc)):
 for _ in range(2):
  _lib.check(fn(), ctx.handle)
 torch.cuda.synchronize()
 e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
 e0.record()
 for _ in range(5):
  fn()
 e1.record(); torch.cuda.synchronize()
 ms = e0.elapsed_time(e1) / 5
 flop = 4.0 * N * D * B
 print(name + ' local gradient pass: %.3f ms for %d chains x %d rows x %d dims = %.1f TFLOP/s algorithmic; X shard %.0f MB -> %.0f GB/s if read once'
       % (ms, B, N, D, flop / ms / 1e9, N * tg.padded_dim * 4 / 1e6, N * tg.padded_dim * 4 / ms / 1e6), flush=True)
k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=2e-3, num_leapfrog_steps=10)
st = th.clone()
kr = k.bootstrap_results(st)
s, kr = k.one_step(st, kr, seed=(1, 2))
torch.cuda.synchronize()
e0.record()
for i in range(3):
  s, kr = k.one_step(s, kr, seed=(3, i))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print('HMC transition (L=10, one rank, no all-reduce): %.2f ms -> %.3e chain-grad/s on this shard; accept %.2f'
      % (ms, B * 10 / ms * 1e3, kr.is_accepted.float().mean().item()))
