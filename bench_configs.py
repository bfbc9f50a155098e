"""Per-config legs of bench.py: the named BASELINE.json configs other than the headline C2, each bounded to a few
seconds of GPU time, reported like C2 -- grad-evals/s, min-ESS/s, roofline, e2e through the public API with host
buffers, and (N = 1 only) the CPU port of the reference algorithm on the box's host cores, min-ESS/s included.

  C1  Eight Schools non-centred, HMC eps=0.4 L=3, 64 chains x 1000 steps                       (N = 1)
  C3  logistic regression 1000 x 25, NUTS depth 10 + DualAveraging IN the timed region,
      8,192 chains per GPU, chains sharded over the ranks (adaptation statistics reduced inside pb2_run)
  C4  stochastic volatility T = 2516, NUTS depth 10, 4,096 chains                               (N = 1)
  C5  logistic regression 1e6 rows x 100 weights, 1,024 replicated chains, HMC L = 10, rows sharded over the ranks,
      per-leapfrog gradient all-reduce enqueued from C (pb2_rowshard_leapfrog)

`shard_parity` (N > 1): before timing, a small chain-sharded NUTS + dual-averaging run and a row-sharded HMC run are
compared with the unsharded runs of the same job on every rank (tests/multigpu_check.py's invariants).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

UNIT = 'grad-evals/s'
FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 148 SMs x 128 FMA lanes x 2 flop x 1.965 GHz = 74.4


def log(*a):
  print(*a, file=sys.stderr, flush=True)


class Env:
  """What every leg needs: device, ranks, peaks."""

  def __init__(self, dev, rank, world, peaks, peaks_src):
    self.dev, self.rank, self.world, self.peaks, self.peaks_src = dev, rank, world, peaks, peaks_src
    self.tf32_peak = 0.5 * peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops'))
    self.hbm_peak = peaks.get('hbm_gbs')


def _timed(fn):
  import torch
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  out = fn()
  e1.record()
  torch.cuda.synchronize()
  return out, e0.elapsed_time(e1) / 1e3


def _max_over_ranks(env, seconds, count):
  """(max seconds over ranks, sum of counts over ranks)"""
  import torch
  import torch.distributed as dist
  if env.world == 1:
    return seconds, count
  t = torch.tensor([seconds], device=env.dev, dtype=torch.float64)
  c = torch.tensor([float(count)], device=env.dev, dtype=torch.float64)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  dist.all_reduce(c, op=dist.ReduceOp.SUM)
  return float(t.item()), float(c.item())


def _flat_draws(draws):
  import torch
  if torch.is_tensor(draws):
    return draws
  return torch.cat([d.reshape(d.shape[0], d.shape[1], -1) for d in draws], -1)


def ess_block(tfp, draws, B, seconds, max_chains=1024):
  """min over dimensions of ESS / s with the reference estimator (diagnostic.py:203-336): cross-chain
  (`cross_chain_dims=1`, the headline `min_ess_per_sec`) and the sum over chains of per-chain ESS.  The latter is
  reported only when it is a valid (positive) number: with short chains the per-chain estimator's
  `-1 + 2 sum rho` can go negative (the estimator's behaviour, identical in the oracle)."""
  flat = _flat_draws(draws)
  nsub = min(B, max_chains)
  sub = flat[:, :nsub].contiguous()
  cross = tfp.mcmc.effective_sample_size(sub, cross_chain_dims=1, filter_beyond_positive_pairs=True,
                                         filter_threshold=None) * (B / nsub)
  per = tfp.mcmc.effective_sample_size(sub, filter_beyond_positive_pairs=True, filter_threshold=None).sum(0) * (B / nsub)
  rhat = tfp.mcmc.potential_scale_reduction(sub, split_chains=True)
  per_min = float(per.min())
  return {'min_ess_per_sec': float(cross.min()) / seconds,
          'min_ess_per_sec_sum_over_chains': per_min / seconds if per_min > 0 else None,
          'max_split_rhat': float(rhat.max()), 'draws_per_chain': int(flat.shape[0]), 'chains_used_for_ess': nsub,
          'sampling_seconds': seconds}


def cpu_port(kind, otgt, x, eps, budget_s, min_draws=4, **kw):
  """The oracle port on the host cores over a bounded sample; grad-evals/s and (cross-chain) min-ESS/s of its draws."""
  import bench
  with bench.all_host_threads():       # every host core, also under torchrun's OMP_NUM_THREADS=1
    return _cpu_port(kind, otgt, x, eps, budget_s, min_draws, bench.blas_threads(), **kw)


def _cpu_port(kind, otgt, x, eps, budget_s, min_draws, threads, **kw):
  from oracle import diagnostic as odiag
  from oracle import mcmc as omcmc
  from oracle import rng as orng
  lp, g = otgt.logp_grad(x)
  seed = orng.sanitize_seed(17, salt='mcmc.sample_chain')
  t0 = time.perf_counter()
  n_grad, done, draws = 0, 0, []
  while (time.perf_counter() - t0) < budget_s or done < min_draws:
    s, seed = orng.split(seed, 2)
    if kind == 'hmc':
      r = omcmc.hmc_one_step(otgt, x, lp, g, eps, kw['L'], s)
      n_grad += x.shape[0] * kw['L']
    else:
      r = omcmc.nuts_one_step(otgt, x, lp, g, eps, s, max_tree_depth=kw['depth'])
      n_grad += int(r['leapfrogs_taken'].sum())
    x, lp, g = r['state'], r['target_log_prob'], r['grads']
    draws.append(x.copy())
    done += 1
    if done >= kw.get('max_steps', 10 ** 9):
      break
  dt = time.perf_counter() - t0
  out = {'value': n_grad / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port',
         'sample': '%d chains x %d %s transitions, NumPy float32 port of the reference algorithm '
                   '(oracle/mcmc.py), BLAS threads = %d of %d host cores; %.1fs' % (x.shape[0], done, kind.upper(), threads,
                                                                                    os.cpu_count(), dt)}
  if done >= 4 and x.shape[0] > 1:
    st = np.stack(draws)
    ess = odiag.effective_sample_size(st, cross_chain_dims=1, filter_beyond_positive_pairs=True, filter_threshold=None)
    mn = float(np.min(ess))
    # too few draws can make the estimator's `-1 + 2 sum rho` non-positive: not a usable number
    out['min_ess_per_sec'] = mn / dt if mn > 0 else None
    out['min_ess_note'] = 'cross-chain ESS of the %d draws x %d chains of this sample' % (done, x.shape[0])
  return out


def e2e_leg(tfp, env, kernel, state_host, steps, grads_of, leapfrogs_per_step=None):
  """The same metric through the public API with HOST buffers: every step copies the state host -> device from
  pinned memory, runs sample_chain(1) and reads the new state (+ leapfrog counts) back."""
  import torch
  was_list = isinstance(state_host, (list, tuple))
  parts = list(state_host) if was_list else [state_host]
  pin_in = [torch.from_numpy(np.ascontiguousarray(p)).pin_memory() for p in parts]
  pin_out = [torch.empty_like(p).pin_memory() for p in pin_in]
  B = pin_in[0].shape[0]
  cnt = torch.empty(B, dtype=torch.int32).pin_memory()
  h2d = sum(p.numel() * 4 for p in pin_in)
  d2h = h2d + (B * 4 if leapfrogs_per_step is None else 0)
  total = 0

  def one(i):
    nonlocal total, pin_in, pin_out
    st = [p.to(env.dev, non_blocking=True) for p in pin_in]
    tr = (lambda _, kr: grads_of(kr)) if leapfrogs_per_step is None else None
    r = tfp.mcmc.sample_chain(1, st if was_list else st[0], kernel=kernel, trace_fn=tr, seed=1000 + i)
    states = r.all_states if tr is not None else r
    outs = list(states) if was_list else [states]
    for o, p in zip(outs, pin_out):
      p.copy_(o[0], non_blocking=True)
    if tr is not None:
      cnt.copy_(r.trace[0], non_blocking=True)
    torch.cuda.synchronize()
    total += int(cnt.sum()) if tr is not None else leapfrogs_per_step * B
    pin_in, pin_out = pin_out, pin_in

  one(-1)   # warm-up (allocations)
  total = 0
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for i in range(steps):
    one(i)
  dt = time.perf_counter() - t0
  dt, tot = _max_over_ranks(env, dt, total)
  return {'value': tot / dt, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': steps,
          'api': 'tfp.mcmc.sample_chain(num_results=1) per step from pinned host buffers'}


# ------------------------------------------------------------------------------------------------ C1
def run_c1(tfp, env, cpu=True):
  import torch
  from oracle import targets as otargets
  tg = tfp.targets.EightSchools()
  B, steps = 64, 1000
  mk = lambda: [torch.zeros(B, device=env.dev), torch.zeros(B, device=env.dev), torch.ones(B, 8, device=env.dev)]
  hmc = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.4, num_leapfrog_steps=3)
  state = [s[0] for s in tfp.mcmc.sample_chain(1, mk(), kernel=hmc, num_burnin_steps=50, trace_fn=None, seed=2)]
  tfp.mcmc.sample_chain(steps, state, kernel=hmc, trace_fn=None, seed=4)
  best, draws = 1e30, None
  for rep in range(3):   # a ~3 ms launch: best of 3
    d, dt = _timed(lambda: tfp.mcmc.sample_chain(steps, state, kernel=hmc, trace_fn=None, seed=3))
    if dt < best:
      best, draws = dt, d
  n = B * steps * 3
  out = {'config': 'C1 Eight Schools non-centred HMC eps=0.4 L=3, 64 chains x 1000 steps', 'chains': B, 'steps': steps,
         'value': n / best, 'unit': UNIT, 'seconds': best,
         'roofline': {'bound': 'latency', 'achieved': best / steps * 1e6, 'unit': 'us per 3-leapfrog transition',
                      'peak': None, 'frac': None,
                      'note': '64 chains = 2 warps on one SM: 3,000 dependent leapfrogs; neither HBM nor tensor bound '
                              '(SURVEY 8d): algorithmic traffic is the 2.56 MB trace'}}
  out.update({k: v for k, v in ess_block(tfp, draws, B, best).items()})
  st_host = [np.zeros(B, np.float32), np.zeros(B, np.float32), np.ones((B, 8), np.float32)]
  out['e2e'] = e2e_leg(tfp, env, hmc, st_host, 20, None, leapfrogs_per_step=3)
  if cpu:
    x0 = np.tile(np.array([0, 0] + [1] * 8, np.float32), (B, 1))
    out['cpu_baseline'] = cpu_port('hmc', otargets.EightSchools(), x0, np.float32(0.4), 3.0, L=3, max_steps=1000)
  return out


# ------------------------------------------------------------------------------------------------ C3
def run_c3(tfp, env, cpu=True, chains=8192, steps=200, adapt=160):
  """NUTS depth 10 + DualAveraging with the ADAPTATION PHASE INSIDE the timed region (80 % of the steps adapt, the
  config's num_adaptation = 0.8 burn-in): one fused sample_chain call; with N > 1 ranks the chains are sharded and the
  accept statistic of every adapting transition is reduced across the ranks inside pb2_run (2-float all-gather)."""
  import torch
  from oracle import targets as otargets
  X, y = otargets.synthetic_logistic_data(1000, 24, seed=0)   # X includes the ones column
  tg = tfp.targets.LogisticRegression(X[:, :-1], y)
  B = chains
  shard = tfp.mcmc.ChainShard(env.rank * B, env.world * B)
  nuts = tfp.mcmc.NoUTurnSampler(tg, step_size=0.1, max_tree_depth=10, experimental_chain_shard=shard)
  k = tfp.mcmc.DualAveragingStepSizeAdaptation(
      nuts, num_adaptation_steps=adapt, experimental_reduce_chain_axis_names='ranks' if env.world > 1 else None)
  st0 = torch.zeros(B, 25, device=env.dev)
  ctx = tfp._lib.Context.get(env.dev)

  def run(seed, tot):
    return tfp.mcmc.sample_chain(steps, st0, kernel=k, trace_fn=None, seed=seed, experimental_leapfrog_total=tot,
                                 return_final_kernel_results=True)
  tot = torch.zeros(B, dtype=torch.int64, device=env.dev)
  run(1, tot)    # warm-up of the same call
  best, n, res = 1e30, 0, None
  l0 = ctx.launch_count()
  for rep in range(2):
    tot.zero_()
    if env.world > 1:
      import torch.distributed as dist
      dist.barrier()
    r, dt = _timed(lambda: run(2, tot))
    dt, cnt = _max_over_ranks(env, dt, float(tot.sum().item()))
    if dt < best:
      best, n, res = dt, cnt, r
  launches = (ctx.launch_count() - l0) // 2
  eps = float(tfp.mcmc.dual_averaging_step_size_adaptation._flat(res.final_kernel_results.new_step_size)[0])
  flops = 4.0 * 1000 * 25 * n
  ach = flops / best / 1e12
  out = {'config': 'C3 logistic regression 1000x25 NUTS depth 10 + DualAveraging (adaptation timed: %d of %d steps adapt), '
                   '%d chains per GPU, chains sharded over %d GPU(s)' % (adapt, steps, B, env.world),
         'chains_per_gpu': B, 'n_gpus': env.world, 'steps': steps, 'value': n / best, 'unit': UNIT, 'seconds': best,
         'gpu_launches': int(launches), 'adapted_step_size': eps,
         'leapfrogs_per_transition': n / (steps * B * env.world),
         'roofline': {'bound': 'tensor', 'achieved': ach / env.world, 'peak': env.tf32_peak, 'unit': 'TFLOP/s',
                      'frac': ach / env.world / env.tf32_peak,
                      'note': 'per GPU; algorithmic 4*N*D = 100,000 flop per chain-gradient against the dense TF32 peak '
                              '(0.5 x measured sustained bf16); this kernel is chain_kernel<WarpG,LogisticT<25>,NUTS> on '
                              'the FP32 pipe (FFMA2), not on tcgen05 -- DESIGN.md section 5'}}
  if env.rank == 0:
    final = res.all_states[-1].contiguous()
    fixed = nuts.copy(step_size=eps, experimental_chain_shard=None)
    nd = 100
    d, dts = _timed(lambda: tfp.mcmc.sample_chain(nd, final[:2048].contiguous(), kernel=fixed, trace_fn=None, seed=5))
    eb = ess_block(tfp, d, 2048, dts)
    eb['note'] = 'ESS from %d post-adaptation draws of 2,048 chains on rank 0 (fixed adapted step size)' % nd
    out.update(eb)
  st_host = np.zeros((B, 25), np.float32)
  out['e2e'] = e2e_leg(tfp, env, nuts.copy(step_size=eps), st_host, 5, lambda kr: kr.leapfrogs_taken)
  if cpu and env.world == 1:
    ot = otargets.LogisticRegression(X, y)
    x0 = (0.1 * np.random.default_rng(2).standard_normal((256, 25))).astype(np.float32)
    out['cpu_baseline'] = cpu_port('nuts', ot, x0, np.float32(eps), 8.0, depth=10)
  return out


# ------------------------------------------------------------------------------------------------ C4
def run_c4(tfp, env, cpu=True, chains=4096, steps=8, adapt=60):
  import torch
  from oracle import targets as otargets
  yret = otargets.synthetic_sv_returns(2516, seed=0)
  tg = tfp.targets.StochasticVolatility(yret)
  B = chains
  st = torch.zeros(B, 2519, device=env.dev)
  nuts = tfp.mcmc.NoUTurnSampler(tg, step_size=0.05, max_tree_depth=10)
  k = tfp.mcmc.DualAveragingStepSizeAdaptation(nuts, num_adaptation_steps=adapt)
  res = tfp.mcmc.sample_chain(1, st, kernel=k, num_burnin_steps=adapt + adapt // 4, trace_fn=None, seed=1,
                              return_final_kernel_results=True)
  eps = float(tfp.mcmc.dual_averaging_step_size_adaptation._flat(res.final_kernel_results.new_step_size)[0])
  st = res.all_states[0].contiguous()
  nuts = nuts.copy(step_size=eps)
  tot = torch.zeros(B, dtype=torch.int64, device=env.dev)
  tfp.mcmc.sample_chain(2, st, kernel=nuts, trace_fn=None, seed=4)
  _, dt = _timed(lambda: tfp.mcmc.sample_chain(steps, st, kernel=nuts, trace_fn=None, seed=3,
                                               experimental_leapfrog_total=tot))
  n = float(tot.sum().item())
  gflop = 63e3 * n / dt / 1e12
  out = {'config': 'C4 stochastic volatility T=2516 (synthetic S&P500-shape returns) NUTS depth 10, %d chains; step size '
                   'from %d untimed dual-averaging steps' % (B, adapt),
         'chains': B, 'steps': steps, 'value': n / dt, 'unit': UNIT, 'seconds': dt, 'adapted_step_size': eps,
         'leapfrogs_per_transition': n / (steps * B),
         'roofline': {'bound': 'fp32', 'achieved': gflop, 'peak': FP32_TFLOPS, 'unit': 'TFLOP/s', 'frac': gflop / FP32_TFLOPS,
                      'hbm_equivalent_gbs': 60456.0 * n / dt / 1e9, 'hbm_peak_gbs': env.hbm_peak,
                      'note': 'the chain state stays on chip (registers + shared memory), so HBM sees only results: the '
                              'kernel is bound by FP32 issue (two block-wide scans of affine maps per gradient, ~63 kflop '
                              '+ T exp); peak = 148 SMs x 128 lanes x 2 flop x 1.965 GHz; hbm_equivalent_gbs = what a '
                              'state-streaming implementation would have to move (60,456 B per leapfrog, SURVEY 8d)'}}
  nd = 40
  d, dts = _timed(lambda: tfp.mcmc.sample_chain(nd, st[:1024].contiguous(), kernel=nuts, trace_fn=None, seed=5))
  out.update(ess_block(tfp, d, 1024, dts))
  out['e2e'] = e2e_leg(tfp, env, nuts, np.ascontiguousarray(st.cpu().numpy()), 3, lambda kr: kr.leapfrogs_taken)
  if cpu:
    ot = otargets.StochasticVolatility(yret)
    out['cpu_baseline'] = cpu_port('nuts', ot, np.ascontiguousarray(st[:32].cpu().numpy()), np.float32(eps), 8.0,
                                   depth=10)
  return out


# ------------------------------------------------------------------------------------------------ C5
C5_ROWS, C5_D, C5_B, C5_L = 1_000_000, 100, 1024, 10


def c5_shard(rank, world):
  """Synthetic 1e6 x 100 design: every rank generates its own contiguous block of rows (same model everywhere)."""
  per = C5_ROWS // world
  rng = np.random.default_rng(1000 + rank)
  theta_true = np.random.default_rng(1).standard_normal(C5_D).astype(np.float32) * 0.1
  X = rng.standard_normal((per, C5_D - 1), dtype=np.float32)
  z = X @ theta_true[:-1] + theta_true[-1]
  y = (rng.random(per) < 1 / (1 + np.exp(-z))).astype(np.float32)
  return X, y


def run_c5(tfp, env, cpu=True, steps=10, tune=30):
  """HMC L=10 with the step size tuned (untimed) to accept ~0.75 by dual averaging through the step loop."""
  import torch
  from oracle import targets as otargets
  if env.world > 1:
    assert tfp.distribute.init_comm() == env.world
  ctx0 = tfp._lib.Context.get(env.dev)
  collective = 'none (one rank)' if env.world == 1 else 'peer-memory reduction fused into the step kernel (CUDA IPC over NVLink)'
  if os.environ.get('PB2_ROWSHARD_COLLECTIVE') == '0' and env.world > 1:
    collective = 'NCCL all-reduce enqueued from C between the kernels'
  X, y = c5_shard(env.rank, env.world)
  tg = tfp.targets.RowShardedLogisticRegression(X, y)
  B, D, L = C5_B, C5_D, C5_L
  hmc = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=1.0e-3, num_leapfrog_steps=L)
  st = torch.zeros(B, D, device=env.dev)
  # posterior sd ~ 2e-3: start the chains near the mode with a few large-acceptance transitions, then adapt
  k = tfp.mcmc.DualAveragingStepSizeAdaptation(hmc, num_adaptation_steps=tune)
  kr = k.bootstrap_results(st)
  for i in range(tune):
    st, kr = k.one_step(st, kr, seed=(7, i))
  eps = float(kr.new_step_size)
  hmc = hmc.copy(step_size=eps)
  kr = hmc.bootstrap_results(st)
  for i in range(3):
    st, kr = hmc.one_step(st, kr, seed=(8, i))
  ctx = tfp._lib.Context.get(env.dev)
  torch.cuda.synchronize()
  if env.world > 1:
    import torch.distributed as dist
    dist.barrier()
  l0 = ctx.launch_count()
  acc = []

  def loop():
    nonlocal st, kr
    for i in range(steps):
      st, kr = hmc.one_step(st, kr, seed=(9, i))
      acc.append(kr.is_accepted)
  _, dt = _timed(loop)
  launches = ctx.launch_count() - l0
  dt, _ = _max_over_ranks(env, dt, 0)
  accept = float(torch.stack(acc).float().mean())
  n = B * L * steps
  flops = 4.0 * C5_ROWS * D * n          # all rows, all ranks
  ach = flops / dt / 1e12
  out = {'config': 'C5 logistic regression %d rows x %d weights, %d replicated chains, HMC L=%d, rows sharded over %d '
                   'GPU(s), per-leapfrog gradient sum over the ranks inside pb2_rowshard_leapfrog' % (C5_ROWS, D, B, L, env.world),
         'n_gpus': env.world, 'chains': B, 'steps': steps, 'value': n / dt, 'unit': UNIT, 'seconds': dt,
         'ms_per_transition': 1e3 * dt / steps, 'step_size': eps, 'accept_rate': accept, 'gpu_launches': int(launches),
         'scaling': 'strong', 'collective': collective,
         'roofline': {'bound': 'tensor', 'achieved': ach / env.world, 'peak': env.tf32_peak, 'unit': 'TFLOP/s',
                      'frac': ach / env.world / env.tf32_peak, 'frac_of_3xtf32_peak': ach / env.world / (env.tf32_peak / 3),
                      'note': 'per GPU; algorithmic 4*N*D = 4e8 flop per chain-gradient; the timed transition includes the momentum '
                              'draw, the cross-rank gradient sum of every leapfrog and the Metropolis-Hastings kernel'}}
  # min-ESS/s: a short run of draws (every transition is ~ms: bounded)
  nd = 24
  draws = []

  def sample():
    nonlocal st, kr
    for i in range(nd):
      st, kr = hmc.one_step(st, kr, seed=(10, i))
      draws.append(st)
  _, dts = _timed(sample)
  out.update(ess_block(tfp, torch.stack(draws), B, dts))
  # e2e: host state in, one transition, host state out (the data shard stays resident: it is the model)
  pin_in = torch.from_numpy(np.ascontiguousarray(st.cpu().numpy())).pin_memory()
  pin_out = torch.empty_like(pin_in).pin_memory()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  es = 5
  for i in range(es):
    s = pin_in.to(env.dev, non_blocking=True)
    s2, _ = hmc.one_step(s, hmc.bootstrap_results(s), seed=(11, i))
    pin_out.copy_(s2, non_blocking=True)
    torch.cuda.synchronize()
    pin_in, pin_out = pin_out, pin_in
  e2e_dt, _ = _max_over_ranks(env, time.perf_counter() - t0, 0)
  out['e2e'] = {'value': B * (L + 1) * es / e2e_dt, 'unit': UNIT, 'h2d_bytes_per_step': B * D * 4,
                'd2h_bytes_per_step': B * D * 4, 'steps': es,
                'api': 'HamiltonianMonteCarlo.bootstrap_results + one_step per step from pinned host buffers (the '
                       'bootstrap gradient is counted)'}
  if cpu and env.world == 1:
    # bounded sample: 65,536 of the rows x 128 chains (the port is BLAS-bound: cost is linear in rows x chains)
    rows, chains = 65536, 128
    Xb = np.concatenate([X[:rows], np.ones((rows, 1), np.float32)], 1)
    ot = otargets.LogisticRegression(Xb, y[:rows])
    x0 = np.ascontiguousarray(st[:chains].cpu().numpy())
    cb = cpu_port('hmc', ot, x0, np.float32(eps), 6.0, L=2, min_draws=2)
    scale = (rows / C5_ROWS)
    cb['value_sample'] = cb['value']
    cb['value'] = cb['value'] * scale
    cb.pop('min_ess_per_sec', None)
    cb['sample'] += '; on %d of the 1e6 rows: value = sample rate x %d/1e6 (cost linear in rows)' % (rows, rows)
    out['cpu_baseline'] = cb
  return out


# ------------------------------------------------------------------------------------------------ shard parity
def shard_parity(tfp, env):
  """sharded == unsharded on every rank, with the library-owned communicator (N > 1).  Returns 'ok' or the error."""
  import torch
  import torch.distributed as dist
  try:
    rank, world, dev = env.rank, env.world, env.dev
    assert tfp.distribute.init_comm() == world
    # (a) chain-sharded NUTS + dual averaging (fused, statistics reduced inside pb2_run) == the whole job on one GPU
    Bg = 64 * world
    B = Bg // world
    tg = tfp.targets.EightSchools()
    rng = np.random.default_rng(0)
    x_all = (np.array([0, 0] + [1] * 8) + 0.3 * rng.standard_normal((Bg, 10))).astype(np.float32)

    def parts(x):
      x = torch.tensor(x, device=dev)
      return [x[:, 0].contiguous(), x[:, 1].contiguous(), x[:, 2:].contiguous()]

    def run(state, shard, axis):
      k = tfp.mcmc.NoUTurnSampler(tg, step_size=0.1, max_tree_depth=5, experimental_chain_shard=shard)
      k = tfp.mcmc.DualAveragingStepSizeAdaptation(k, num_adaptation_steps=6, experimental_reduce_chain_axis_names=axis)
      return tfp.mcmc.sample_chain(8, state, kernel=k, seed=11,
                                   trace_fn=lambda _, kr: (kr.inner_results.step_size, kr.inner_results.leapfrogs_taken))
    sh = run(parts(x_all[rank * B:(rank + 1) * B]), tfp.mcmc.ChainShard(rank * B, Bg), 'ranks')
    fu = run(parts(x_all), None, None)
    # the cross-rank accept statistic is an exact (fixed-point) sum: the adapted step sizes agree bit for bit
    np.testing.assert_array_equal(sh.trace[0].cpu().numpy(), fu.trace[0].cpu().numpy())
    np.testing.assert_array_equal(sh.trace[1].cpu().numpy(), fu.trace[1][:, rank * B:(rank + 1) * B].cpu().numpy())
    for a, b in zip(sh.all_states, fu.all_states):
      np.testing.assert_array_equal(a.cpu().numpy(), b[:, rank * B:(rank + 1) * B].cpu().numpy())
    # (b) row-sharded HMC (tcgen05 gradient, all-reduce inside pb2_rowshard_leapfrog) == all rows on one GPU; replicas
    # take bit-identical decisions
    n, d, Bc = 4096, 39, 256
    X, y = tfp.targets.synthetic_logistic_data(n, d, seed=1)
    per = n // world
    lo, hi = rank * per, (n if rank == world - 1 else (rank + 1) * per)
    t_shard = tfp.targets.RowShardedLogisticRegression(X[lo:hi], y[lo:hi])
    t_full = tfp.targets.RowShardedLogisticRegression(X, y)
    t_full._world = lambda: None
    st = torch.tensor((0.1 * np.random.default_rng(2).standard_normal((Bc, d + 1))).astype(np.float32), device=dev)
    outs = []
    for t in (t_shard, t_full):
      k = tfp.mcmc.HamiltonianMonteCarlo(t, step_size=0.01, num_leapfrog_steps=4)
      outs.append(tfp.mcmc.sample_chain(4, st, kernel=k, seed=5, trace_fn=lambda _, kr: (kr.is_accepted, kr.log_accept_ratio)))
    s, f = outs
    np.testing.assert_allclose(s.trace[1].cpu().numpy(), f.trace[1].cpu().numpy(), rtol=5e-3, atol=5e-3)
    assert (s.trace[0] == f.trace[0]).float().mean().item() > 0.97
    allst = [torch.zeros_like(s.all_states) for _ in range(world)]
    dist.all_gather(allst, s.all_states.contiguous())
    for a in allst:
      assert torch.equal(a, s.all_states), 'replicated chains differ between ranks'
    ok = torch.ones(1, device=dev)
  except Exception as e:  # pylint: disable=broad-except
    msg = '%s: %s' % (type(e).__name__, str(e)[:300])
    ok = torch.zeros(1, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return msg
  dist.all_reduce(ok, op=dist.ReduceOp.MIN)
  return 'ok' if float(ok.item()) == 1.0 else 'failed on another rank'
