"""where does one end-to-end step (host buffers) of bench.py's e2e loop spend its time?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
rng = np.random.default_rng(0)
L = np.linalg.cholesky(tg.covariance)
B, D = 16384, 100
x_pinned = torch.from_numpy((rng.standard_normal((B, D)) @ L.T).astype(np.float32)).pin_memory()
out_pinned = torch.empty(B, D).pin_memory(); cnt_pinned = torch.empty(B, dtype=torch.int32).pin_memory()
nuts = tfp.mcmc.NoUTurnSampler(tg, step_size=0.7415, max_tree_depth=10)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
acc = np.zeros(5)
for i in range(12):
  t0 = sync()
  st = x_pinned.to(dev, non_blocking=True); t1 = sync()
  kr = nuts.bootstrap_results(st); t2 = sync()
  r = tfp.mcmc.sample_chain(1, st, kernel=nuts, previous_kernel_results=kr, trace_fn=lambda _, k: k.leapfrogs_taken, seed=100 + i); t3 = sync()
  out_pinned.copy_(r.all_states[0], non_blocking=True); cnt_pinned.copy_(r.trace[0], non_blocking=True); t4 = sync()
  n = int(cnt_pinned.sum()); x_pinned, out_pinned = out_pinned, x_pinned; t5 = time.perf_counter()
  if i >= 2: acc += np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4])
print('ms per step: H2D %.3f  bootstrap %.3f  sample_chain(1) %.3f  D2H %.3f  host sum/swap %.3f  total %.3f' % (*(acc / 10 * 1e3), acc.sum() / 10 * 1e3))
# kernel-only time of the transition
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = x_pinned.to(dev); kr = nuts.bootstrap_results(st)
e0.record(); s2, k2 = nuts.one_step(st, kr, seed=(1, 2)); e1.record(); torch.cuda.synchronize()
print('one_step (events) %.3f ms' % e0.elapsed_time(e1))
