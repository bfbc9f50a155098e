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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
st = torch.tensor((rng.standard_normal((B, 100)) @ L.T).astype(np.float32), device=dev)
k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=10)
out = {}
for v in (1, 3, 0):
  ctx.set_int('dense_variant', v)
  res = tfp.mcmc.sample_chain(8, st, kernel=k, seed=2, trace_fn=lambda _, kr: (kr.leapfrogs_taken, kr.has_divergence, kr.energy, kr.target_log_prob, kr.is_accepted, kr.log_accept_ratio))
  out[v] = [t.cpu().numpy() for t in res.trace] + [res.all_states.cpu().numpy()]
for v in (1, 3, 0):
  nl = out[v][0]
  print('variant', v, 'count nl==1:', (nl == 1).sum(0).nonzero()[0][:20], 'n chains', ((nl == 1).sum(0) > 0).sum(), 'divergences', out[v][1].sum())
bad = ((out[0][0] == 1).sum(0) > 0).nonzero()[0]
for c in bad[:6]:
  for v in (1, 3, 0):
    print('chain', c, 'variant', v, 'nl', out[v][0][:, c], 'div', out[v][1][:, c].astype(int), 'lp', np.round(out[v][3][:, c], 2), 'en', np.round(out[v][2][:, c], 2), 'lar', np.round(out[v][5][:, c], 3))
