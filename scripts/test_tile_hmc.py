"""tile (tcgen05) HMC vs warp-per-chain HMC vs oracle on the 100-d ill-conditioned Gaussian."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import probability_b200 as tfp
from probability_b200 import _lib
from oracle import mcmc as omcmc, rng as orng, targets as otargets

dev = torch.device('cuda', 0)
tg = tfp.targets.IllConditionedGaussian()
ctx = _lib.Context.get(dev)
B = 300
rng = np.random.default_rng(0)
L = np.linalg.cholesky(tg.covariance)
x0 = (rng.standard_normal((B, 100)) @ L.T).astype(np.float32)
st = torch.tensor(x0, device=dev)
k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.5, num_leapfrog_steps=5)
seed = orng.key(4)
outs = {}
for variant in (0, 1):
  ctx.set_int('dense_variant', variant)
  s, r = k.one_step(st, k.bootstrap_results(st), seed=seed)
  torch.cuda.synchronize()
  outs[variant] = (s.cpu().numpy(), r.is_accepted.cpu().numpy(), r.log_accept_ratio.cpu().numpy(),
                   r.proposed_state.cpu().numpy(), r.proposed_results.initial_momentum.cpu().numpy())
o32 = otargets.DenseGaussian(tg.precision, tg.log_normalizer)
lp0, g0 = o32.logp_grad(x0)
ref = omcmc.hmc_one_step(o32, x0, lp0, g0, 0.5, 5, seed)
for v in (0, 1):
  s, acc, lar, prop, m0 = outs[v]
  print('variant', v, 'accept agree', (acc == ref['is_accepted']).mean(), 'lar maxdiff', np.nanmax(np.abs(lar - ref['log_accept_ratio'])),
        'prop rel', np.max(np.abs(prop - ref['proposed_state'])) / np.abs(ref['proposed_state']).max(),
        'm0 maxdiff', np.abs(m0 - ref['initial_momentum']).max(), 'acc rate', acc.mean(), flush=True)
print('tile vs warp: accept agree', (outs[0][1] == outs[1][1]).mean(), 'state maxdiff', np.abs(outs[0][0] - outs[1][0]).max())
# throughput: HMC L=32, 16384 chains
B = 16384
x0 = (rng.standard_normal((B, 100)) @ L.T).astype(np.float32)
st = torch.tensor(x0, device=dev)
k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.5, num_leapfrog_steps=32)
for variant in (0, 1):
  ctx.set_int('dense_variant', variant)
  tfp.mcmc.sample_chain(2, st, kernel=k, trace_fn=None, seed=1)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  res = tfp.mcmc.sample_chain(20, st, kernel=k, trace_fn=lambda _, kr: kr.is_accepted, seed=2)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  print('variant %d: 20 transitions x %d chains x L=32 in %.2f ms -> %.3e grad-evals/s, accept %.3f' % (
      variant, B, ms, 20 * B * 32 / ms * 1e3, res.trace.float().mean().item()), flush=True)
