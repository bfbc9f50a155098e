"""A few fused NUTS launches on the stochastic-volatility target (C4 shape) for ncu: the NUTS chain kernel is the longest
launch of the run."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
dev = torch.device('cuda', 0)
y = tfp.targets.synthetic_sv_returns(2516, seed=0)
tg = tfp.targets.StochasticVolatility(y)
B = 1184
st = torch.zeros(B, 2519, device=dev)
k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.04, max_tree_depth=10)
for rep in range(3):
  r = tfp.mcmc.sample_chain(2, st, kernel=k, trace_fn=lambda _, kr: kr.leapfrogs_taken, seed=rep)
  torch.cuda.synchronize()
print('mean leapfrogs', float(r.trace.float().mean()))
