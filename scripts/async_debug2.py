import sys
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
ctx = _lib.Context.get(dev)
rng = np.random.default_rng(4)
L = np.linalg.cholesky(tg.covariance)
B = 700
x0 = torch.tensor((rng.standard_normal((B, 100)) @ L.T).astype(np.float32), device=dev)
k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.7, max_tree_depth=8)
for variant in (0, 3):
  ctx.set_int('dense_variant', variant)
  fused = tfp.mcmc.sample_chain(4, x0, kernel=k, trace_fn=lambda _, kr: (kr.leapfrogs_taken, kr.target_log_prob), seed=11)
  loop = tfp.mcmc.sample_chain(4, x0, kernel=k, trace_fn=lambda _, kr: (kr.leapfrogs_taken + 0, kr.target_log_prob), seed=11)
  for tt in range(4):
    ds = (loop.all_states[tt] != fused.all_states[tt]).any(-1).sum().item()
    dl = (loop.trace[0][tt] != fused.trace[0][tt]).sum().item()
    dp = (loop.trace[1][tt] != fused.trace[1][tt]).sum().item()
    print('variant', variant, 'transition', tt, 'chains differing: states', ds, 'leapfrogs', dl, 'lp', dp)
