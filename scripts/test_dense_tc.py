"""Standalone check of the tcgen05 dense-Gaussian gradient against float64 (run under `timeout`)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib

dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
ctx = _lib.Context.get(dev); ctx.bind_stream()
for B in (128, 300, 16384):
  rng = np.random.default_rng(B)
  L = np.linalg.cholesky(tg.covariance)
  x = (rng.standard_normal((B, 100)) @ L.T).astype(np.float32)
  xt = torch.tensor(x, device=dev)
  lp = torch.empty(B, device=dev); g = torch.empty(B, 100, device=dev)
  _lib.check(ctx.lib.pb2_dense_logp_grad_tc(ctx.handle, tg.handle(ctx), B, _lib.ptr(xt), _lib.ptr(lp), _lib.ptr(g)), ctx.handle)
  torch.cuda.synchronize()
  P64 = tg.precision.astype(np.float64)
  g64 = -(x.astype(np.float64) @ P64)
  lp64 = 0.5 * np.sum(x * g64, 1) + tg.log_normalizer
  lp2, g2 = tg.log_prob_and_grad(xt)      # FFMA path
  scale = np.abs(g64).max(1, keepdims=True)
  e_tc = np.max(np.abs(g.cpu().numpy() - g64) / scale)
  e_ff = np.max(np.abs(g2.cpu().numpy() - g64) / scale)
  print('B=%d grad err tc %.2e ffma %.2e | lp err tc %.2e ffma %.2e' % (
      B, e_tc, e_ff, np.max(np.abs(lp.cpu().numpy() - lp64) / np.abs(lp64)),
      np.max(np.abs(lp2.cpu().numpy() - lp64) / np.abs(lp64))), flush=True)
# timing
B = 16384 * 8
xt = torch.randn(B, 100, device=dev)
lp = torch.empty(B, device=dev); g = torch.empty(B, 100, device=dev)
for name, fn in (('tc', lambda: ctx.lib.pb2_dense_logp_grad_tc(ctx.handle, tg.handle(ctx), B, _lib.ptr(xt), _lib.ptr(lp), _lib.ptr(g))),
                 ('ffma', lambda: ctx.lib.pb2_logp_grad(ctx.handle, tg.handle(ctx), B, _lib.ptr(xt), _lib.ptr(lp), _lib.ptr(g)))):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10): fn()
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 10
  print('%s: %.3f ms per %d-chain gradient -> %.3e chain-grad/s, %.1f TFLOP/s algorithmic, %.0f GB/s' % (
      name, ms, B, B / ms * 1e3, 2e4 * B / ms * 1e3 / 1e12, (B * 100 * 4 * 2 + B * 4) / ms * 1e3 / 1e9), flush=True)
