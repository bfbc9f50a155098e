"""small asynchronous-lane NUTS + tile HMC + logistic tensor-core runs, meant for compute-sanitizer"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
from oracle import targets as otargets
dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
rng = np.random.default_rng(0)
L = np.linalg.cholesky(tg.covariance)
x0 = torch.tensor((rng.standard_normal((300, 100)) @ L.T).astype(np.float32), device=dev)
k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.74, max_tree_depth=7)
r = tfp.mcmc.sample_chain(2, x0, kernel=k, trace_fn=lambda _, kr: kr.leapfrogs_taken, seed=1)
print('nuts async ok', float(r.trace.float().mean()))
h = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.5, num_leapfrog_steps=4)
r = tfp.mcmc.sample_chain(2, x0, kernel=h, trace_fn=lambda _, kr: kr.is_accepted, seed=1)
print('hmc tile ok', float(r.trace.float().mean()))
X, y = otargets.synthetic_logistic_data(200, 24, seed=0)
tl = tfp.targets.LogisticRegression(X[:, :-1], y)
ctx = _lib.Context.get(dev); ctx.bind_stream()
th = torch.randn(256, 25, device=dev) * 0.3
lp = torch.empty(256, device=dev); g = torch.empty(256, 25, device=dev)
_lib.check(ctx.lib.pb2_logistic_logp_grad_tc(ctx.handle, tl.handle(ctx), 256, _lib.ptr(th), _lib.ptr(lp), _lib.ptr(g)), ctx.handle)
torch.cuda.synchronize()
print('logistic tc ok', float(lp.mean()))
