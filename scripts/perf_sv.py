"""C4 kernel A/B: stochastic volatility T=2516, NUTS depth 10, 4,096 chains, fixed step size."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
dev = torch.device('cuda', 0)
y = tfp.targets.synthetic_sv_returns(2516, seed=0)
tg = tfp.targets.StochasticVolatility(y)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
st = torch.zeros(B, 2519, device=dev)
k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.05, max_tree_depth=10)
r = tfp.mcmc.sample_chain(1, st, num_burnin_steps=75, trace_fn=None, seed=1, return_final_kernel_results=True,
                          kernel=tfp.mcmc.DualAveragingStepSizeAdaptation(k, num_adaptation_steps=60))
st = r.all_states[0].contiguous()
k = k.copy(step_size=float(r.final_kernel_results.new_step_size))
best = 1e9
for rep in range(2):
  tot = torch.zeros(B, dtype=torch.int64, device=dev)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  tfp.mcmc.sample_chain(8, st, kernel=k, trace_fn=None, seed=2, experimental_leapfrog_total=tot)
  e1.record(); torch.cuda.synchronize()
  best = min(best, e0.elapsed_time(e1))
print('SV NUTS %d chains: %.1f ms, %.3e grad-evals/s (mean leapfrogs %.1f)' % (B, best, tot.sum().item() / best * 1e3, tot.float().mean().item() / 8))
