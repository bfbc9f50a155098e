"""small asynchronous-lane NUTS + tile HMC + logistic tensor-core (TMA-staged, both shapes) + run-time-compiled user
target runs, meant for compute-sanitizer"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
from oracle import targets as otargets
dev = torch.device("cuda", 0)
torch.manual_seed(0)
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

# row-sharded gradient (LargeD shape, TMA tensor copies into the operand ring)
Xr, yr = otargets.synthetic_logistic_data(700, 99, seed=1)
tr = tfp.targets.RowShardedLogisticRegression(Xr[:, :-1], yr)
thr = torch.randn(200, 100, device=dev) * 0.1
lpr, gr = tr.log_prob_and_grad(thr)
torch.cuda.synchronize()
print('rowshard tc ok', float(lpr.mean()))
# user-defined target (NVRTC build of the chain kernels)
src = """
__device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data) {
  float lp = 0.f;
  for (int d = 0; d < 5; ++d) { g[d] = -x[d]; lp -= 0.5f * x[d] * x[d]; }
  return lp;
}
"""
tu = tfp.targets.UserTarget(5, src)
ku = tfp.mcmc.NoUTurnSampler(tu, step_size=0.5, max_tree_depth=4)
ru = tfp.mcmc.sample_chain(3, torch.zeros(64, 5, device=dev), kernel=ku, trace_fn=lambda _, kr: kr.leapfrogs_taken, seed=3)
torch.cuda.synchronize()
print('user target ok', float(ru.trace.float().mean()))
# stochastic volatility (CTA per chain, two CTAs per SM)
yv = tfp.targets.synthetic_sv_returns(200, seed=0)
ts = tfp.targets.StochasticVolatility(yv)
ks = tfp.mcmc.NoUTurnSampler(ts, step_size=0.05, max_tree_depth=4)
rs = tfp.mcmc.sample_chain(2, torch.zeros(40, 203, device=dev), kernel=ks, trace_fn=lambda _, kr: kr.leapfrogs_taken, seed=4)
torch.cuda.synchronize()
print('sv ok', float(rs.trace.float().mean()))
