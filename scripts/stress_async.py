"""stress: asynchronous-lane tile NUTS == lock-step tile NUTS (bit for bit) over batch sizes / depths / step sizes /
numbers of fused transitions, including step sizes that diverge.  Run under `timeout`; prints each case."""
import itertools, sys
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
ctx = _lib.Context.get(dev)
L = np.linalg.cholesky(tg.covariance)
bad = 0
for B, depth, K, eps in itertools.product((256, 257, 1000, 4097, 20000), (6, 8, 10), (1, 2, 7), (0.1, 0.74, 3.0)):
  if B == 20000 and (eps == 0.1 and depth == 10):
    continue   # 1023-leaf trees for 20,000 chains in lock-step: slow, covered at smaller B
  rng = np.random.default_rng(B + depth)
  x0 = torch.tensor((rng.standard_normal((B, 100)) @ L.T).astype(np.float32), device=dev)
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=eps, max_tree_depth=depth)
  out = {}
  for v in (0, 3):
    ctx.set_int('dense_variant', v)
    r = tfp.mcmc.sample_chain(K, x0, kernel=k, trace_fn=lambda _, kr: (kr.leapfrogs_taken, kr.has_divergence, kr.energy), seed=B + K)
    out[v] = [r.all_states.cpu().numpy()] + [f.cpu().numpy() for f in r.trace]
  ctx.set_int('dense_variant', 0)
  same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(out[0], out[3]))
  bad += not same
  print('B=%5d depth=%2d K=%d eps=%.2f: %s  (mean leapfrogs %.1f, divergent %.3f)' % (
      B, depth, K, eps, 'identical' if same else 'DIFFERENT', out[0][1].mean(), out[0][2].mean()), flush=True)
print('cases different:', bad)
