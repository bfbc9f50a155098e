"""One launch of the SV leapfrog kernel (L = 33) for ncu."""
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
lp, g = tg.log_prob_and_grad(xt)
m = torch.randn(B, 2519, device=dev)
step = torch.tensor([1e-3], device=dev)
outs = [torch.empty_like(xt) for _ in range(3)] + [torch.empty(B, device=dev)]
for _ in range(2):
  _lib.check(ctx.lib.pb2_leapfrog(ctx.handle, tg.handle(ctx), B, _lib.ptr(m), _lib.ptr(xt), _lib.ptr(lp), _lib.ptr(g),
                                  _lib.ptr(step), 0, 33, _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[3]), _lib.ptr(outs[2])), ctx.handle)
torch.cuda.synchronize()
