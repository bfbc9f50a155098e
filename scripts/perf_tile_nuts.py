"""async-lane tile NUTS (variant 0) vs lock-step tile (3) vs warp kernels (1): same chains, same trees."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
ctx = _lib.Context.get(dev)
rng = np.random.default_rng(0)
L = np.linalg.cholesky(tg.covariance)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 10
x0 = (rng.standard_normal((B, 100)) @ L.T).astype(np.float32)
st = torch.tensor(x0, device=dev)
k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=depth)
out = {}
import os
variants = (0,) if os.environ.get('PB2_ONLY_SCHED') else (1, 3, 0)
for variant in variants:
  ctx.set_int('dense_variant', variant)
  tfp.mcmc.sample_chain(2, st, kernel=k, trace_fn=None, seed=1)
  torch.cuda.synchronize()
  best = 1e9
  for rep in range(2):
    tot = torch.zeros(B, dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = tfp.mcmc.sample_chain(K, st, kernel=k, trace_fn=lambda _, kr: (kr.leapfrogs_taken, kr.is_accepted, kr.energy), seed=2,
                                experimental_leapfrog_total=tot)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
  out[variant] = (res.trace[0].cpu().numpy(), res.all_states.cpu().numpy(), res.trace[2].cpu().numpy())
  print('variant %d: %d transitions x %d chains: best %.2f ms -> %.3e grad-evals/s (mean leapfrogs %.1f)' % (
      variant, K, B, best, tot.sum().item() / best * 1e3, out[variant][0].mean()), flush=True)
for a, b in (() if os.environ.get('PB2_ONLY_SCHED') else ((0, 3), (0, 1), (3, 1))):
  same = out[a][0] == out[b][0]
  # chains whose whole history of tree sizes agrees
  ok = same.all(0)
  close = np.isclose(out[a][1][-1], out[b][1][-1], rtol=5e-3, atol=5e-2).all(-1)
  print("   final states close: %.4f" % close.mean())
  print('variant %d vs %d: leapfrogs equal %.4f; chains with identical history %.4f; max |state diff| on those %.3e' % (
      a, b, same.mean(), ok.mean(), np.abs(out[a][1][:, ok] - out[b][1][:, ok]).max()), flush=True)
