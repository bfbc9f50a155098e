import sys
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
rng = np.random.default_rng(0)
L = np.linalg.cholesky(tg.covariance)
B = 16384
st = torch.tensor((rng.standard_normal((B, 100)) @ L.T).astype(np.float32), device=dev)
k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=10)
tot = torch.zeros(B, dtype=torch.int64, device=dev)
res = tfp.mcmc.sample_chain(30, st, kernel=k, trace_fn=lambda _, kr: (kr.leapfrogs_taken, kr.is_accepted), seed=2, experimental_leapfrog_total=tot)
x = res.all_states
var = x.var(0)                       # [B, D]
stuck = (var.max(1).values == 0)
print('chains with zero variance over 30 draws:', int(stuck.sum()), 'min leapfrogs', int(res.trace[0].min()), 'nan states', int(torch.isnan(x).sum()))
print('accept mean', res.trace[1].float().mean().item(), 'total leapfrogs == trace sum', int(tot.sum()), int(res.trace[0].sum()))
same_as_prev = (x[1:] == x[:-1]).all(-1).float().mean().item()
print('fraction of (draw,chain) identical to previous draw:', same_as_prev)
ess = tfp.mcmc.effective_sample_size(x[:, :2048].contiguous(), filter_beyond_positive_pairs=True, filter_threshold=None)
print('nan ess entries', int(torch.isnan(ess).sum()), 'min ess', float(ess[~torch.isnan(ess)].min()))
