"""async tile NUTS vs lock-step tile NUTS: where do they differ first?"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
ctx = _lib.Context.get(dev)
rng = np.random.default_rng(3)
L = np.linalg.cholesky(tg.covariance)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 7
x0 = (rng.standard_normal((B, 100)) @ L.T).astype(np.float32)
st = torch.tensor(x0, device=dev)
k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=depth)
fields = lambda _, kr: (kr.leapfrogs_taken, kr.is_accepted, kr.energy, kr.target_log_prob, kr.log_accept_ratio)
out = {}
for name, v in (('a', 0), ('b', 0), ('l', 3)):
  ctx.set_int('dense_variant', v)
  res = tfp.mcmc.sample_chain(K, st, kernel=k, trace_fn=fields, seed=7)
  out[name] = [res.all_states.cpu().numpy()] + [f.cpu().numpy() for f in res.trace]
names = ['states', 'leapfrogs', 'accepted', 'energy', 'lp', 'lar']
for a, b in (('a', 'b'), ('a', 'l')):
  print('---', a, 'vs', b)
  for tt in range(K):
    msg = []
    for nm, u, v in zip(names, out[a], out[b]):
      d = (u[tt] != v[tt])
      if nm == 'states': d = d.any(-1)
      msg.append('%s %d' % (nm, d.sum()))
    print('transition', tt, ' differing chains:', ', '.join(msg))
lf = out['l'][1][0]; la = out['a'][1][0]
bad = np.nonzero((out['a'][0][0] != out['l'][0][0]).any(-1))[0]
print('first bad chains', bad[:10], 'leapfrogs lock', lf[bad[:10]], 'async', la[bad[:10]])
good = np.nonzero(~(out['a'][0][0] != out['l'][0][0]).any(-1))[0]
print('some good chains', good[:10], 'leapfrogs', lf[good[:10]])
import collections
print('bad by leapfrogs', sorted(collections.Counter(lf[bad]).items())[:20])
print('good by leapfrogs', sorted(collections.Counter(lf[good]).items())[:20])
