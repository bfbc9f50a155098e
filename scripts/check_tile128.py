"""128-chain tile NUTS kernels: lock-step (variant 3) vs oracle / warp kernels (1), async (0) == lock-step bit for bit,
then throughput of the new async kernel (0) next to the 64-chain async kernel (4).  Stage = argv[1]."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
from oracle import mcmc as omcmc, rng as orng, targets as otargets

stage = sys.argv[1] if len(sys.argv) > 1 else 'lock'
dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
ctx = _lib.Context.get(dev)
rng = np.random.default_rng(0)
L = np.linalg.cholesky(tg.covariance)


def state(B, seed=0):
  r = np.random.default_rng(seed)
  return (r.standard_normal((B, 100)) @ L.T).astype(np.float32)


if stage == 'lock':
  B = 300
  x0 = state(B)
  st = torch.tensor(x0, device=dev)
  o32 = otargets.DenseGaussian(tg.precision, tg.log_normalizer)
  lp0, g0 = o32.logp_grad(x0)
  for depth, eps in ((3, 0.3), (6, 0.5)):
    k = tfp.mcmc.NoUTurnSampler(tg, step_size=eps, max_tree_depth=depth)
    seed = orng.key(4)
    ref = omcmc.nuts_one_step(o32, x0, lp0, g0, eps, seed, max_tree_depth=depth)
    for variant in (3, 0, 1):
      ctx.set_int('dense_variant', variant)
      s, r = k.one_step(st, k.bootstrap_results(st), seed=seed)
      torch.cuda.synchronize()
      nl = r.leapfrogs_taken.cpu().numpy()
      same = nl == ref['leapfrogs_taken']
      close = np.isclose(s.cpu().numpy(), ref['state'], rtol=2e-3, atol=2e-3).all(1)
      print('depth', depth, 'variant', variant, 'leapfrogs same', same.mean(), 'states close (of same)', close[same].mean(),
            'acc agree', (r.is_accepted.cpu().numpy() == ref['is_accepted'])[same].mean(), 'mean nl', nl.mean(), flush=True)
elif stage == 'async':
  for B, depth, K in ((1000, 7, 3), (4096, 10, 3), (20000, 10, 2)):
    x0 = state(B, 3)
    st = torch.tensor(x0, device=dev)
    k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=depth)
    fields = lambda _, kr: (kr.leapfrogs_taken, kr.is_accepted, kr.energy, kr.target_log_prob, kr.log_accept_ratio)
    outs = {}
    for name, variant in (('async', 0), ('lockstep', 3), ('async_again', 0)):
      ctx.set_int('dense_variant', variant)
      res = tfp.mcmc.sample_chain(K, st, kernel=k, trace_fn=fields, seed=7)
      outs[name] = [res.all_states.cpu().numpy()] + [f.cpu().numpy() for f in res.trace]
    for other in ('lockstep', 'async_again'):
      eq = [np.array_equal(a, b) for a, b in zip(outs['async'], outs[other])]
      print('B', B, 'depth', depth, 'async vs', other, 'bit-identical:', eq, 'leapfrogs equal frac',
            (outs['async'][1] == outs[other][1]).mean(), 'mean leapfrogs', outs['async'][1].mean(), flush=True)
elif stage == 'perf':
  B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
  K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
  variants = [int(v) for v in sys.argv[4].split(',')] if len(sys.argv) > 4 else [0, 4]
  x0 = state(B)
  st = torch.tensor(x0, device=dev)
  k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=10)
  for variant in variants:
    ctx.set_int('dense_variant', variant)
    tfp.mcmc.sample_chain(2, st, kernel=k, trace_fn=None, seed=1)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
      tot = torch.zeros(B, dtype=torch.int64, device=dev)
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      tfp.mcmc.sample_chain(K, st, kernel=k, trace_fn=None, seed=2, experimental_leapfrog_total=tot)
      e1.record(); torch.cuda.synchronize()
      best = min(best, e0.elapsed_time(e1))
    print('variant %d: %d transitions x %d chains: best %.2f ms -> %.3e grad-evals/s (mean leapfrogs %.1f, max chain total %d)' % (
        variant, K, B, best, tot.sum().item() / best * 1e3, tot.float().mean().item() / K, tot.max().item()), flush=True)
