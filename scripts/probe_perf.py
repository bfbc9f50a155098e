"""Quick throughput probe of the named configs (development aid, not the bench)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import probability_b200 as tfp  # noqa: E402


def timed(fn):
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  out = fn()
  e1.record()
  torch.cuda.synchronize()
  return out, e0.elapsed_time(e1) / 1e3


def run(name, target, state, kernel, warm, steps, adapt=None):
  dev = state.device if torch.is_tensor(state) else state[0].device
  B = (state if torch.is_tensor(state) else state[0]).shape[0]
  if adapt:
    k = tfp.mcmc.DualAveragingStepSizeAdaptation(kernel, num_adaptation_steps=adapt)
    tot = torch.zeros(B, dtype=torch.int64, device=dev)
    res, dt = timed(lambda: tfp.mcmc.sample_chain(
        1, state, kernel=k, num_burnin_steps=adapt, trace_fn=None, seed=1, return_final_kernel_results=True,
        experimental_leapfrog_total=tot))
    eps = float(res.final_kernel_results.new_step_size)
    print('%s: adapt %d steps in %.3fs, eps=%.4f, grad evals %.3e -> %.3e/s' % (
        name, adapt, dt, eps, tot.sum().item(), tot.sum().item() / dt), flush=True)
    state = res.all_states[0] if torch.is_tensor(res.all_states) else [s[0] for s in res.all_states]
    kernel = kernel.copy(step_size=eps)
  tot = torch.zeros(B, dtype=torch.int64, device=dev)
  tfp.mcmc.sample_chain(1, state, kernel=kernel, num_burnin_steps=warm, trace_fn=None, seed=2)
  tot.zero_()
  res, dt = timed(lambda: tfp.mcmc.sample_chain(
      steps, state, kernel=kernel, trace_fn=None, seed=3, experimental_leapfrog_total=tot))
  n = tot.sum().item()
  print('%s: %d transitions x %d chains in %.4fs: %.3e grad evals -> %.3e grad-evals/s (%.1f leapfrogs/transition)'
        % (name, steps, B, dt, n, n / dt, n / steps / B), flush=True)


def main():
  dev = torch.device('cuda', 0)
  import os
  which = [a for a in sys.argv[1:] if a.startswith('c')] or ['c1', 'c2', 'c3', 'c4']
  noadapt = os.environ.get('PROBE_NOADAPT') == '1'   # profiling: skip adaptation, use known step sizes
  if 'c1' in which:
    tg = tfp.targets.EightSchools()
    st = [torch.zeros(64, device=dev), torch.zeros(64, device=dev), torch.ones(64, 8, device=dev)]
    run('C1 eight-schools HMC B=64', tg, st, tfp.mcmc.HamiltonianMonteCarlo(tg, 0.4, 3), 10, 1000)
    st = [torch.zeros(65536, device=dev), torch.zeros(65536, device=dev), torch.ones(65536, 8, device=dev)]
    run('C1b eight-schools HMC B=65536', tg, st, tfp.mcmc.HamiltonianMonteCarlo(tg, 0.4, 3), 10, 1000)
    run('C1c eight-schools NUTS B=65536', tg, st, tfp.mcmc.NoUTurnSampler(tg, 0.3, max_tree_depth=10), 3, 20)
  if 'c2' in which:
    tg = tfp.targets.IllConditionedGaussian()
    st = torch.zeros(16384, 100, device=dev)
    run('C2 dense-gaussian NUTS B=16384', tg, st, tfp.mcmc.NoUTurnSampler(tg, 0.74 if noadapt else 0.158, max_tree_depth=10), 2,
        10, adapt=None if noadapt else 60)
  if 'c3' in which:
    X, y = tfp.targets.synthetic_logistic_data(1000, 24, seed=0)
    tg = tfp.targets.LogisticRegression(X, y)
    st = torch.zeros(8192, 25, device=dev)
    run('C3 logistic NUTS B=8192', tg, st, tfp.mcmc.NoUTurnSampler(tg, 0.093 if noadapt else 0.1, max_tree_depth=10),
        2, 10, adapt=None if noadapt else 60)
  if 'c4' in which:
    yv = tfp.targets.synthetic_sv_returns()
    tg = tfp.targets.StochasticVolatility(yv)
    st = torch.zeros(4096, 2519, device=dev)
    st[:, 1] = float(np.log(yv.var()))
    run('C4 stoch-vol NUTS B=4096', tg, st, tfp.mcmc.NoUTurnSampler(tg, 0.038 if noadapt else 0.02, max_tree_depth=10),
        1, 3, adapt=None if noadapt else 30)


if __name__ == '__main__':
  main()
