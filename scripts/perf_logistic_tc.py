"""throughput of the logistic log-prob + gradient primitive: tcgen05 kernel vs the FP32 warp-per-chain kernel"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
from oracle import targets as otargets
dev = torch.device('cuda', 0)
X, y = otargets.synthetic_logistic_data(1000, 24, seed=0)
tg = tfp.targets.LogisticRegression(X[:, :-1], y)
ctx = _lib.Context.get(dev); ctx.bind_stream()
for B in (8192, 65536):
  th = torch.randn(B, 25, device=dev) * 0.3
  lp = torch.empty(B, device=dev); g = torch.empty(B, 25, device=dev)
  for name, fn in (('tcgen05', ctx.lib.pb2_logistic_logp_grad_tc), ('fp32 warp', ctx.lib.pb2_logp_grad)):
    for _ in range(3):
      _lib.check(fn(ctx.handle, tg.handle(ctx), B, _lib.ptr(th), _lib.ptr(lp), _lib.ptr(g)), ctx.handle)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
      fn(ctx.handle, tg.handle(ctx), B, _lib.ptr(th), _lib.ptr(lp), _lib.ptr(g))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('B=%6d %-10s %.3f ms per batched gradient -> %.3e chain-grad/s (%.1f TFLOP/s algorithmic)' % (
        B, name, ms, B / ms * 1e3, B / ms * 1e3 * 1e5 / 1e12), flush=True)
