"""Where a C4 leaf goes: gradient alone, leapfrog alone, NUTS leaf (per-SM latency per chain-gradient)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
dev = torch.device('cuda', 0)
y = tfp.targets.synthetic_sv_returns(2516, seed=0)
tg = tfp.targets.StochasticVolatility(y)
B = 148 * 16
rng = np.random.default_rng(0)
x = (0.05 * rng.standard_normal((B, 2519))).astype(np.float32); x[:, 0] += 2; x[:, 1] += 5
xt = torch.tensor(x, device=dev)
ctx = _lib.Context.get(dev); ctx.bind_stream()
def timed(fn, n=3):
  best = 1e9
  for _ in range(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
  return best
lp, g = tg.log_prob_and_grad(xt)
ms = timed(lambda: tg.log_prob_and_grad(xt))
print('logp_grad: %.3f ms for %d chains -> %.0f cycles per chain-gradient per SM' % (ms, B, ms * 1e-3 * 1.965e9 / (B / 148)))
m = torch.randn(B, 2519, device=dev)
step = torch.tensor([1e-3], device=dev)
outs = [torch.empty_like(xt) for _ in range(3)] + [torch.empty(B, device=dev)]
for L in (1, 33):
  f = lambda: _lib.check(ctx.lib.pb2_leapfrog(ctx.handle, tg.handle(ctx), B, _lib.ptr(m), _lib.ptr(xt), _lib.ptr(lp), _lib.ptr(g),
                                             _lib.ptr(step), 0, L, _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[3]), _lib.ptr(outs[2])), ctx.handle)
  f(); msL = timed(f)
  print('leapfrog L=%d: %.3f ms' % (L, msL))
  if L == 1: ms1 = msL
print('   -> %.0f cycles per leapfrog per SM' % ((msL - ms1) / 32 * 1e-3 * 1.965e9 / (B / 148)))
