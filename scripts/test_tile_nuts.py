"""tile (tcgen05) NUTS vs warp-per-chain NUTS vs oracle on the 100-d ill-conditioned Gaussian."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
from oracle import mcmc as omcmc, rng as orng, targets as otargets

dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
ctx = _lib.Context.get(dev)
rng = np.random.default_rng(0)
L = np.linalg.cholesky(tg.covariance)
if 'perf' not in sys.argv:
  B = 300
  x0 = (rng.standard_normal((B, 100)) @ L.T).astype(np.float32)
  st = torch.tensor(x0, device=dev)
  for depth, eps in ((3, 0.3), (6, 0.5)):
    k = tfp.mcmc.NoUTurnSampler(tg, step_size=eps, max_tree_depth=depth)
    seed = orng.key(4)
    outs = {}
    for variant in (0, 1):
      ctx.set_int('dense_variant', variant)
      s, r = k.one_step(st, k.bootstrap_results(st), seed=seed)
      torch.cuda.synchronize()
      outs[variant] = dict(state=s.cpu().numpy(), nl=r.leapfrogs_taken.cpu().numpy(), lar=r.log_accept_ratio.cpu().numpy(),
                           acc=r.is_accepted.cpu().numpy(), en=r.energy.cpu().numpy(), rm=r.reach_max_depth.cpu().numpy(),
                           dv=r.has_divergence.cpu().numpy(), lp=r.target_log_prob.cpu().numpy())
    o32 = otargets.DenseGaussian(tg.precision, tg.log_normalizer)
    lp0, g0 = o32.logp_grad(x0)
    ref = omcmc.nuts_one_step(o32, x0, lp0, g0, eps, seed, max_tree_depth=depth)
    for v in (0, 1):
      o = outs[v]
      same = o['nl'] == ref['leapfrogs_taken']
      close = np.isclose(o['state'], ref['state'], rtol=2e-3, atol=2e-3).all(1)
      print('depth', depth, 'variant', v, 'leapfrogs same', same.mean(), 'states close (of same)', close[same].mean(),
            'acc agree', (o['acc'] == ref['is_accepted'])[same].mean(), 'lar maxdiff', np.nanmax(np.abs(o['lar'] - ref['log_accept_ratio'])[same & close]),
            'mean nl', o['nl'].mean(), flush=True)
    print('   tile vs warp: leapfrogs same', (outs[0]['nl'] == outs[1]['nl']).mean(), flush=True)
# throughput
B = 16384
x0 = (rng.standard_normal((B, 100)) @ L.T).astype(np.float32)
st = torch.tensor(x0, device=dev)
k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=10)
for variant in (0, 1):
  ctx.set_int('dense_variant', variant)
  tfp.mcmc.sample_chain(1, st, kernel=k, trace_fn=None, seed=1)
  torch.cuda.synchronize()
  tot = torch.zeros(B, dtype=torch.int64, device=dev)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  res = tfp.mcmc.sample_chain(5, st, kernel=k, trace_fn=lambda _, kr: kr.leapfrogs_taken, seed=2, experimental_leapfrog_total=tot)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  nl = res.trace.float()
  print('variant %d: 5 transitions x %d chains in %.2f ms -> %.3e grad-evals/s; leapfrogs mean %.1f max %d; tile util %.3f' % (
      variant, B, ms, tot.sum().item() / ms * 1e3, nl.mean().item(), int(nl.max().item()),
      (nl.reshape(5, -1, 128).mean(-1) / nl.reshape(5, -1, 128).max(-1).values).mean().item()), flush=True)
