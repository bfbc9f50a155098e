"""phase profile of the lock-step tile NUTS kernel (build with PB2_NVCC_EXTRA=-DPB2_TILE_PROF)."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
ctx = _lib.Context.get(dev)
rng = np.random.default_rng(0)
L = np.linalg.cholesky(tg.covariance)
B = 16384
x0 = (rng.standard_normal((B, 100)) @ L.T).astype(np.float32)
st = torch.tensor(x0, device=dev)
k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=10)
ctx.set_int('dense_variant', 3)
for rep in range(2):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  tot = torch.zeros(B, dtype=torch.int64, device=dev)
  e0.record()
  tfp.mcmc.sample_chain(3, st, kernel=k, trace_fn=None, seed=2, experimental_leapfrog_total=tot)
  e1.record(); torch.cuda.synchronize()
  print('%.2f ms, %.3e grad-evals/s' % (e0.elapsed_time(e1), tot.sum().item() / e0.elapsed_time(e1) * 1e3))
