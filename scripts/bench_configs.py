"""The named configs other than bench.py's C2, each reported like bench.py reports C2: leapfrog gradient
evaluations/s and min-ESS/s on one B200 next to the CPU oracle port on the box's host cores (same run).

  C1  Eight Schools non-centred, HMC eps=0.4 L=3, 64 chains x 1000 steps  (+ the same at 65,536 chains)
  C3  Bayesian logistic regression 1000 x 25 (synthetic German-credit shape), NUTS depth 10 + DualAveraging,
      8,192 chains per GPU
  C4  stochastic volatility (synthetic series, T=2516), NUTS depth 10 + DualAveraging, 4,096 chains

usage: python scripts/bench_configs.py [c1 c3 c4] > profiles/r01_configs.jsonl   (one JSON line per config)
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probability_b200 as tfp  # noqa: E402
from oracle import mcmc as omcmc  # noqa: E402  (CPU baseline leg only)
from oracle import rng as orng  # noqa: E402
from oracle import targets as otargets  # noqa: E402

DEV = torch.device('cuda', 0)


def log(*a):
  print(*a, file=sys.stderr, flush=True)


def timed(fn):
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  out = fn()
  e1.record()
  torch.cuda.synchronize()
  return out, e0.elapsed_time(e1) / 1e3


def ess_block(draws, B, seconds):
  """min over dimensions of ESS / s: sum over chains (default threshold) and cross-chain."""
  flat = draws if torch.is_tensor(draws) else torch.cat([d.reshape(d.shape[0], d.shape[1], -1) for d in draws], -1)
  nsub = min(B, 1024)
  sub = flat[:, :nsub].contiguous()
  per = tfp.mcmc.effective_sample_size(sub, filter_beyond_positive_pairs=True, filter_threshold=None).sum(0) * (B / nsub)
  cross = tfp.mcmc.effective_sample_size(sub, cross_chain_dims=1, filter_beyond_positive_pairs=True,
                                         filter_threshold=None) * (B / nsub)
  rhat = tfp.mcmc.potential_scale_reduction(sub, split_chains=True)
  return {'min_ess_per_sec_sum_over_chains': float(per.min()) / seconds,
          'min_ess_per_sec_cross_chain': float(cross.min()) / seconds, 'max_split_rhat': float(rhat.max()),
          'draws_per_chain': int(flat.shape[0]), 'chains_used_for_ess': nsub}


def gpu_run(name, state, kernel, steps, adapt, draws):
  B = (state if torch.is_tensor(state) else state[0]).shape[0]
  eps = None
  if adapt:
    k = tfp.mcmc.DualAveragingStepSizeAdaptation(kernel, num_adaptation_steps=adapt)
    res = tfp.mcmc.sample_chain(1, state, kernel=k, num_burnin_steps=adapt + adapt // 4, trace_fn=None, seed=1,
                                return_final_kernel_results=True)
    eps = float(res.final_kernel_results.new_step_size)
    state = res.all_states[0] if torch.is_tensor(res.all_states) else [s[0] for s in res.all_states]
    kernel = kernel.copy(step_size=eps)
  else:
    state = tfp.mcmc.sample_chain(1, state, kernel=kernel, num_burnin_steps=50, trace_fn=None, seed=2)
    state = state[0] if torch.is_tensor(state) else [s[0] for s in state]
  tot = torch.zeros(B, dtype=torch.int64, device=DEV)
  tfp.mcmc.sample_chain(3, state, kernel=kernel, trace_fn=None, seed=4)   # warm-up of the fused driver
  dt = 1e30
  for _ in range(3):   # best of 3: the small configs are a few milliseconds long (host overhead, allocator state)
    tot.zero_()
    _, d1 = timed(lambda: tfp.mcmc.sample_chain(steps, state, kernel=kernel, trace_fn=None, seed=3,
                                                 experimental_leapfrog_total=tot))
    dt = min(dt, d1)
  n = float(tot.sum().item())
  out = {'config': name, 'chains': B, 'steps': steps, 'seconds': dt, 'value': n / dt, 'unit': 'grad-evals/s',
         'leapfrogs_per_transition': n / steps / B, 'step_size': eps}
  d, dts = timed(lambda: tfp.mcmc.sample_chain(draws, state, kernel=kernel, trace_fn=None, seed=5))
  out['min_ess'] = ess_block(d, B, dts)
  log('%s: %.3e grad-evals/s (%.1f leapfrogs/transition, eps %s)' % (name, out['value'], out['leapfrogs_per_transition'], eps))
  return out


def cpu_run(kind, otgt, x, eps, budget_s, **kw):
  """oracle port on the host cores (BLAS threads = all of them, also under torchrun's OMP_NUM_THREADS=1), bounded sample"""
  try:
    from threadpoolctl import threadpool_limits, threadpool_info
  except ImportError:
    return _cpu_run(kind, otgt, x, eps, budget_s, os.cpu_count(), **kw)
  with threadpool_limits(limits=os.cpu_count()):
    threads = max([d.get('num_threads', 1) for d in threadpool_info()] or [1])
    return _cpu_run(kind, otgt, x, eps, budget_s, threads, **kw)


def _cpu_run(kind, otgt, x, eps, budget_s, threads, **kw):
  lp, g = otgt.logp_grad(x)
  seed = orng.sanitize_seed(17, salt='mcmc.sample_chain')
  t0 = time.perf_counter()
  n_grad, done = 0, 0
  while (time.perf_counter() - t0) < budget_s:
    s, seed = orng.split(seed, 2)
    if kind == 'hmc':
      r = omcmc.hmc_one_step(otgt, x, lp, g, eps, kw['L'], s)
      n_grad += x.shape[0] * kw['L']
    else:
      r = omcmc.nuts_one_step(otgt, x, lp, g, eps, s, max_tree_depth=kw['depth'])
      n_grad += int(r['leapfrogs_taken'].sum())
    x, lp, g = r['state'], r['target_log_prob'], r['grads']
    done += 1
  dt = time.perf_counter() - t0
  return {'value': n_grad / dt, 'unit': 'grad-evals/s', 'cores': threads, 'kind': 'port',
          'sample': '%d chains x %d %s transitions, NumPy float32 port of the reference algorithm; %.1fs' % (
              x.shape[0], done, kind.upper(), dt)}


def main():
  which = [a for a in sys.argv[1:] if a.startswith('c')] or ['c1', 'c3', 'c4']
  if 'c1' in which:
    tg = tfp.targets.EightSchools()
    mk = lambda B: [torch.zeros(B, device=DEV), torch.zeros(B, device=DEV), torch.ones(B, 8, device=DEV)]
    hmc = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.4, num_leapfrog_steps=3)
    o = gpu_run('C1 Eight Schools HMC eps=0.4 L=3, 64 chains x 1000 steps', mk(64), hmc, 1000, 0, 1000)
    x0 = np.tile(np.array([0, 0] + [1] * 8, np.float32), (64, 1))
    o['cpu_baseline'] = cpu_run('hmc', otargets.EightSchools(), x0, np.float32(0.4), 10.0, L=3)
    print(json.dumps(o), flush=True)
    o = gpu_run('C1b Eight Schools HMC eps=0.4 L=3, 65,536 chains x 1000 steps', mk(65536), hmc, 1000, 0, 200)
    print(json.dumps(o), flush=True)
  if 'c3' in which:
    X, y = otargets.synthetic_logistic_data(1000, 24, seed=0)   # X includes the ones column
    tg = tfp.targets.LogisticRegression(X[:, :-1], y)
    B = 8192
    st = torch.zeros(B, 25, device=DEV)
    nuts = tfp.mcmc.NoUTurnSampler(tg, step_size=0.1, max_tree_depth=10)
    o = gpu_run('C3 logistic regression 1000x25 NUTS depth 10 + DualAveraging, 8,192 chains', st, nuts, 30, 120, 200)
    ot = otargets.LogisticRegression(X, y)
    x0 = (0.1 * np.random.default_rng(2).standard_normal((512, 25))).astype(np.float32)
    o['cpu_baseline'] = cpu_run('nuts', ot, x0, np.float32(o['step_size']), 15.0, depth=10)
    print(json.dumps(o), flush=True)
  if 'c4' in which:
    yret = otargets.synthetic_sv_returns(2516, seed=0)
    tg = tfp.targets.StochasticVolatility(yret)
    B = 4096
    st = torch.zeros(B, 2519, device=DEV)
    nuts = tfp.mcmc.NoUTurnSampler(tg, step_size=0.05, max_tree_depth=10)
    o = gpu_run('C4 stochastic volatility T=2516 NUTS depth 10 + DualAveraging, 4,096 chains', st, nuts, 10, 100, 60)
    ot = otargets.StochasticVolatility(yret)
    x0 = np.zeros((64, 2519), np.float32)
    o['cpu_baseline'] = cpu_run('nuts', ot, x0, np.float32(o['step_size']), 20.0, depth=10)
    print(json.dumps(o), flush=True)


if __name__ == '__main__':
  main()
