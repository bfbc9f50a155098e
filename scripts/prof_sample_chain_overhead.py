"""host-side cost of one tfp.mcmc.sample_chain(num_results=1) call (C1: Eight Schools HMC, 64 chains): cProfile top."""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
dev = torch.device('cuda', 0)
tg = tfp.targets.EightSchools()
k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.4, num_leapfrog_steps=3)
x = torch.zeros(64, 10, device=dev)
h = torch.zeros(64, 10).pin_memory()
for i in range(20):
  tfp.mcmc.sample_chain(1, x, kernel=k, seed=i, trace_fn=None)
torch.cuda.synchronize()
def step(i):
  xd = h.to(dev, non_blocking=True)
  r = tfp.mcmc.sample_chain(1, xd, kernel=k, seed=i, trace_fn=None)
  h.copy_(r[0] if r.dim() == 3 else r)
t = time.perf_counter()
for i in range(200): step(i)
print('per step %.1f us' % ((time.perf_counter() - t) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for i in range(200): step(i)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
