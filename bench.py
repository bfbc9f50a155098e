#!/usr/bin/env python
"""bench.py -- headline benchmark of the many-chain NUTS hot path (BASELINE.json).

Workload (N=1): configs[1] of BASELINE.json -- ill-conditioned dense Gaussian, 100-d, NUTS
(max_tree_depth 10) over 16,384 chains per GPU; chains start from exact target draws, the
step size comes from an untimed dual-averaging warm-up.  A "step" is one NUTS transition of
every chain.  metric = leapfrog gradient evaluations per second (sum of leapfrogs_taken over
chains and timed transitions / device time, max over ranks).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, libpb2)
  python bench.py --impl reference [...]                        CPU arm: the oracle port of the
                                                                TFP algorithm on the host cores
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of tile_nuts_async_kernel per NUTS transition of 16,384 chains: the r02 ncu
# capture of a timed 20-transition launch (0.91 GB read + 9.68 GB written) / 20
NCU_DRAM_BYTES_PER_TRANSITION = 5.3e8
METRIC = 'leapfrog_grad_evals_per_sec'
UNIT = 'grad-evals/s'
D = 100
MAX_DEPTH = 10
CHAINS_PER_GPU = 16384
ADAPT_STEPS = 150
EPS0 = 0.158          # 0.5 * D**-0.25 (windowed_sampling.py:566-589)


def log(*a):
  print(*a, file=sys.stderr, flush=True)


def peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    d = json.load(open(p))
    return d, 'measured'
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


def exact_draws(cov, n, seed):
  rng = np.random.default_rng(seed)
  L = np.linalg.cholesky(cov)
  return (rng.standard_normal((n, cov.shape[0])) @ L.T).astype(np.float32)


class ClockSampler:
  """nvidia-smi clocks / throttle reasons DURING the timed region."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index = index
    self.proc = None
    self.lines = []

  def start(self):
    if os.environ.get('PB2_BENCH_NO_SMI'):   # diagnosis only: does the poller perturb the measurement?
      return
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
           '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.th = threading.Thread(target=self._read, daemon=True)
      self.th.start()
    except Exception:  # pylint: disable=broad-except
      self.proc = None

  def _read(self):
    for ln in self.proc.stdout:
      self.lines.append((time.time(), ln.strip()))

  def stop(self, t_begin=None, t_end=None):
    """Samples that arrived in [t_begin, t_end + one period] (the poller is started long before the timed region:
    the first nvidia-smi of a fresh box takes hundreds of milliseconds to initialise and stalls the GPU meanwhile)."""
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    sm, mx, reasons = [], [], set()
    lines = self.lines
    if t_begin is not None:
      inside = [(ts, ln) for ts, ln in lines if t_begin <= ts <= t_end + 0.12]
      lines = inside if inside else lines[-2:]   # a region shorter than the polling period: the nearest samples
    for ts, ln in lines:
      f = [x.strip() for x in ln.split(',')]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); mx.append(float(f[2]))
      except ValueError:
        continue
      for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
        if v.lower().startswith('active'):
          reasons.add(name)
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(max(mx)) if mx else None,
            'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------
def all_host_threads():
  """BLAS threads = every host core for the CPU legs (torchrun exports OMP_NUM_THREADS=1 to its ranks)."""
  try:
    from threadpoolctl import threadpool_limits
    return threadpool_limits(limits=os.cpu_count())
  except ImportError:
    import contextlib
    return contextlib.nullcontext()


def blas_threads():
  try:
    from threadpoolctl import threadpool_info
    return max([d.get('num_threads', 1) for d in threadpool_info()] or [1])
  except ImportError:
    return int(os.environ.get('OMP_NUM_THREADS', os.cpu_count()))


def cpu_reference_arm(steps, warmup, budget_s=25.0):
  """The oracle port (lock-step batched NumPy restatement of nuts.py) on the host cores."""
  with all_host_threads():
    return _cpu_reference_arm(steps, warmup, budget_s)


def _cpu_reference_arm(steps, warmup, budget_s):
  from oracle import mcmc as omcmc
  from oracle import rng as orng
  from oracle import targets as otargets
  cov, _ = otargets.ill_conditioned_covariance(D)
  P, c = otargets.gaussian_precision_from_cov(cov)
  tgt = otargets.DenseGaussian(P, c)
  B = 1024
  x = exact_draws(cov, B, seed=123)
  lp, g = tgt.logp_grad(x)
  eps = np.float32(0.7)   # close to what dual averaging finds on the GPU arm
  seed = orng.sanitize_seed(17, salt='mcmc.sample_chain')
  for _ in range(max(1, min(warmup, 1))):
    s, seed = orng.split(seed, 2)
    r = omcmc.nuts_one_step(tgt, x, lp, g, eps, s, max_tree_depth=MAX_DEPTH)
    x, lp, g = r['state'], r['target_log_prob'], r['grads']
  t0 = time.perf_counter()
  n_grad = 0
  done = 0
  while done < max(1, steps) and (time.perf_counter() - t0) < budget_s:
    s, seed = orng.split(seed, 2)
    r = omcmc.nuts_one_step(tgt, x, lp, g, eps, s, max_tree_depth=MAX_DEPTH)
    x, lp, g = r['state'], r['target_log_prob'], r['grads']
    n_grad += int(r['leapfrogs_taken'].sum())
    done += 1
  dt = time.perf_counter() - t0
  return {'value': n_grad / dt, 'unit': UNIT, 'cores': blas_threads(), 'kind': 'port',
          'sample': '%d chains x %d NUTS transitions (depth<=%d, eps=%.2f), NumPy float32 lock-step port of '
                    'tfp nuts.py, BLAS threads = %d of %d host cores; %.1fs' % (B, done, MAX_DEPTH, eps, blas_threads(),
                                                                               os.cpu_count(), dt)}, done, dt


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=100)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours')
  ap.add_argument('--chains', type=int, default=CHAINS_PER_GPU)
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-ess', action='store_true')
  ap.add_argument('--step-size', type=float, default=None, help='skip dual averaging (profiling runs)')
  ap.add_argument('--configs', default='all',
                  help="per_config legs: 'all' (N=1: c1,c3,c4,c5; N>1: the sharded configs c3,c5), 'none', or a "
                       "comma list of c1,c3,c4,c5")
  args = ap.parse_args()
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  config = {'workload': 'ill-conditioned dense Gaussian 100-d (inference_gym seed 10), NUTS max_tree_depth=10, '
                        '%d chains per GPU' % args.chains,
            'chains_per_gpu': args.chains, 'chains_total': args.chains * max(world, 1), 'dim': D,
            'max_tree_depth': MAX_DEPTH, 'parallelism': 'chains sharded, no data-path collective',
            'l2': 'flushed (256 MiB write) before every timed launch; the K timed steps are ONE fused sample_chain launch; '
                  'the K-step region is measured 3 times and the median repetition is reported (timed_reps_ms)',
            'init': 'exact target draws; step size from %d untimed dual-averaging steps' % ADAPT_STEPS}

  if args.impl == 'reference':
    if rank != 0:
      return
    cb, done, dt = cpu_reference_arm(args.steps, args.warmup)
    # the CPU arm runs a bounded SAMPLE of the workload: say so in its own config instead of echoing the GPU arm's
    config = dict(config, chains_per_gpu=1024, chains_total=1024,
                  workload=config['workload'].replace('%d chains per GPU' % args.chains,
                                                      'bounded CPU sample: 1,024 of the %d chains' % args.chains),
                  parallelism='host cores (BLAS threads)', l2='n/a (CPU)',
                  init='exact target draws; fixed step size 0.7 (what dual averaging finds on the GPU arm)')
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': done, 'warmup': min(args.warmup, 1), 'ms_per_step': 1e3 * dt / max(done, 1),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': config, 'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))
    return

  import torch
  import torch.distributed as dist
  import probability_b200 as tfp
  from probability_b200 import _lib
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device (B200); there is no CPU fallback for the product path.')
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  shard_parity = None
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
    # the library owns the on-path collectives (NCCL communicator attached to the context); before any timing, the
    # sharded runs are checked against the unsharded job on every rank
    import bench_configs
    pk0, pk0_src = peaks()
    shard_parity = bench_configs.shard_parity(tfp, bench_configs.Env(dev, rank, world, pk0, pk0_src))
    log('[rank %d] shard_parity: %s' % (rank, shard_parity))
  B = args.chains
  shard = tfp.mcmc.ChainShard(chain_offset=rank * B, num_chains_global=world * B)
  target = tfp.targets.IllConditionedGaussian(ndims=D)
  x_host = exact_draws(target.covariance, world * B, seed=123)[rank * B:(rank + 1) * B]
  x_pinned = torch.from_numpy(np.ascontiguousarray(x_host)).pin_memory()
  state = x_pinned.to(dev, non_blocking=True)
  ctx = _lib.Context.get(dev)

  # ---- untimed: dual-averaging warm-up (cross-rank log-mean-exp when sharded)
  nuts = tfp.mcmc.NoUTurnSampler(target, step_size=EPS0, max_tree_depth=MAX_DEPTH, experimental_chain_shard=shard)
  da = tfp.mcmc.DualAveragingStepSizeAdaptation(
      nuts, num_adaptation_steps=ADAPT_STEPS,
      experimental_reduce_chain_axis_names='ranks' if world > 1 else None)
  t0 = time.perf_counter()
  if args.step_size is None:
    res = tfp.mcmc.sample_chain(1, state, kernel=da, num_burnin_steps=ADAPT_STEPS, trace_fn=None, seed=17,
                                return_final_kernel_results=True)
    torch.cuda.synchronize()
    eps = float(res.final_kernel_results.new_step_size)
    state = res.all_states[0].contiguous()
  else:  # profiling runs: skip the adaptation launches
    eps = args.step_size
  log('[rank %d] warm-up %.2fs, adapted step size %.4f' % (rank, time.perf_counter() - t0, eps))
  nuts = nuts.copy(step_size=eps)
  pkr = nuts.bootstrap_results(state)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
  seed = tfp.random.sanitize_seed(18, salt='mcmc.sample_chain')

  def one_transition(st, kr, sd):
    step_seed, sd = tfp.random.split_seed(sd)
    st, kr = nuts.one_step(st, kr, seed=step_seed)
    return st, kr, sd

  sampler = ClockSampler(local_rank)   # polls during warm-up too; only the timed region's samples are reported
  sampler.start()
  for _ in range(max(args.warmup, 3)):
    state, pkr, seed = one_transition(state, pkr, seed)
  # warm-up of the fused driver itself (same code path and the same K as the timed launch: allocations, module load)
  wres = tfp.mcmc.sample_chain(max(args.warmup, 3, args.steps), state, kernel=nuts, previous_kernel_results=pkr,
                               trace_fn=None, seed=20, return_final_kernel_results=True)
  state, pkr = wres.all_states[-1].contiguous(), wres.final_kernel_results
  del wres
  torch.cuda.synchronize()
  # the first nvidia-smi query of a fresh box takes hundreds of milliseconds and stalls the GPU meanwhile: let it finish
  # before the timed region (bounded wait; the library calls no longer block the host, so the warm-up is short)
  t_wait = time.time()
  while sampler.proc is not None and not sampler.lines and time.time() - t_wait < 3.0:
    time.sleep(0.02)
  # ... and bring the clocks back up after that idle wait: one more untimed launch of the timed shape
  wres = tfp.mcmc.sample_chain(max(args.warmup, 3, args.steps), state, kernel=nuts, previous_kernel_results=pkr,
                               trace_fn=None, seed=21, return_final_kernel_results=True)
  state, pkr = wres.all_states[-1].contiguous(), wres.final_kernel_results
  del wres
  torch.cuda.synchronize()

  # ---- timed region: K transitions of every chain as ONE sample_chain call (the fused driver: key schedule
  # kernel + persistent transition kernel), L2 flushed before it, CUDA events on the launch stream.  The region is
  # measured 3 times (fresh seeds, L2 flushed each time) and the MEDIAN repetition is reported: a fresh box's first
  # seconds are noisy (all three are listed in `timed_reps_ms`).
  reps = []
  t_begin = time.time()
  launches0 = ctx.launch_count()
  for rep in range(3):
    flush.fill_(1)
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    tot = torch.zeros(B, dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = tfp.mcmc.sample_chain(args.steps, state, kernel=nuts, previous_kernel_results=pkr, trace_fn=None,
                                seed=19 + 1000 * rep, experimental_leapfrog_total=tot, return_final_kernel_results=True)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    tt = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
    ng = torch.tensor([float(tot.sum().item())], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(tt, op=dist.ReduceOp.MAX)
      dist.all_reduce(ng, op=dist.ReduceOp.SUM)
    reps.append((float(tt.item()), float(ng.item()), float(tot.sum().item())))
    state = res.all_states[-1].contiguous()
    pkr = res.final_kernel_results
    del res
  launches = (ctx.launch_count() - launches0) // 3
  t_end = time.time()
  clocks = sampler.stop(t_begin, t_end)
  total_s, n_grad_all, n_grad = sorted(reps)[1]
  value = n_grad_all / total_s
  step_ms = [1e3 * total_s]     # one launch covers the K steps

  # ---- the same K transitions as K separate one_step launches (what adaptation / TransitionKernel.one_step
  # users see), L2 flushed between launches
  evs, leap = [], []
  for _ in range(args.steps):
    flush.fill_(1)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    state, pkr, seed = one_transition(state, pkr, seed)
    a1.record()
    evs.append((a0, a1))
    leap.append(pkr.leapfrogs_taken)
  torch.cuda.synchronize()
  pl_s = sum(a.elapsed_time(b) for a, b in evs) / 1e3
  pl = torch.tensor([float(sum(int(l.sum().item()) for l in leap)), pl_s], device=dev, dtype=torch.float64)
  if world > 1:
    g = [torch.zeros_like(pl) for _ in range(world)]
    dist.all_gather(g, pl)
    fused_value = sum(float(v[0]) for v in g) / max(float(v[1]) for v in g)
  else:
    fused_value = float(pl[0]) / float(pl[1])

  # ---- e2e: the public API with HOST buffers, per step: H2D state, sample_chain(1), D2H state + counts
  out_pinned = torch.empty(B, D, dtype=torch.float32).pin_memory()
  cnt_pinned = torch.empty(B, dtype=torch.int32).pin_memory()
  x_pinned.copy_(state.cpu())
  e2e_steps = max(3, min(args.steps, 10))
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  e2e_grad = 0
  for i in range(e2e_steps):
    st = x_pinned.to(dev, non_blocking=True)
    r = tfp.mcmc.sample_chain(1, st, kernel=nuts, trace_fn=lambda _, kr: kr.leapfrogs_taken, seed=100 + i)
    out_pinned.copy_(r.all_states[0], non_blocking=True)
    cnt_pinned.copy_(r.trace[0], non_blocking=True)
    torch.cuda.synchronize()
    e2e_grad += int(cnt_pinned.sum())
    x_pinned, out_pinned = out_pinned, x_pinned   # the step's host output is the next step's host input
  e2e_s = time.perf_counter() - t0
  ee = torch.tensor([float(e2e_grad), e2e_s], device=dev, dtype=torch.float64)
  if world > 1:
    g = [torch.zeros_like(ee) for _ in range(world)]
    dist.all_gather(g, ee)
    e2e_value = sum(float(v[0]) for v in g) / max(float(v[1]) for v in g)
  else:
    e2e_value = e2e_grad / e2e_s

  # ---- the same, as ONE user call: H2D initial state, sample_chain(num_results=K), D2H of all K states
  all_pinned = torch.empty(e2e_steps, B, D, dtype=torch.float32).pin_memory()
  tot1 = torch.zeros(B, dtype=torch.int64, device=dev)
  for rep in range(2):   # the first call of this shape is a warm-up (allocations)
    tot1.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = x_pinned.to(dev, non_blocking=True)
    r = tfp.mcmc.sample_chain(e2e_steps, st, kernel=nuts, trace_fn=None, seed=300 + rep, experimental_leapfrog_total=tot1)
    all_pinned.copy_(r, non_blocking=True)
    torch.cuda.synchronize()
    one_call_s = time.perf_counter() - t0
  oc = torch.tensor([float(tot1.sum().item()), one_call_s], device=dev, dtype=torch.float64)
  if world > 1:
    g = [torch.zeros_like(oc) for _ in range(world)]
    dist.all_gather(g, oc)
    one_call_value = sum(float(v[0]) for v in g) / max(float(v[1]) for v in g)
  else:
    one_call_value = float(oc[0]) / float(oc[1])
  del r

  # ---- min-ESS/s (second half of the metric): 200 traced draws per chain
  ess_info = None
  if not args.no_ess and rank == 0:
    n_draws = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    draws = tfp.mcmc.sample_chain(n_draws, state, kernel=nuts, previous_kernel_results=pkr, trace_fn=None, seed=21)
    e1.record()
    torch.cuda.synchronize()
    samp_s = e0.elapsed_time(e1) / 1e3
    sub = draws[:, :2048].contiguous()
    ess = tfp.mcmc.effective_sample_size(sub, filter_beyond_positive_pairs=True, filter_threshold=None)
    per_dim = ess.sum(0) * (B / 2048.0)
    cross = tfp.mcmc.effective_sample_size(sub, cross_chain_dims=1, filter_beyond_positive_pairs=True,
                                           filter_threshold=None) * (B / 2048.0)
    rhat = tfp.mcmc.potential_scale_reduction(sub, split_chains=True)
    mean_err = float((draws.mean((0, 1)).abs() / torch.tensor(np.sqrt(np.diag(target.covariance)),
                                                               device=dev, dtype=torch.float32)).max())
    ess_info = {'min_ess_per_sec_sum_over_chains': float(per_dim.min()) / samp_s,
                'min_ess_per_sec_cross_chain': float(cross.min()) / samp_s, 'draws_per_chain': n_draws,
                'sampling_seconds': samp_s, 'max_split_rhat': float(rhat.max()),
                'max_abs_mean_over_sd': mean_err,
                'note': 'ESS on 2,048 of the chains scaled to all chains of rank 0'}
    del draws, sub

  # ---- the other named configs (bounded), every rank takes part in the sharded ones
  per_config = {}
  if args.configs != 'none':
    import bench_configs
    pk1, pk1_src = peaks()
    env = bench_configs.Env(dev, rank, world, pk1, pk1_src)
    want = (['c1', 'c3', 'c4', 'c5'] if world == 1 else ['c3', 'c5']) if args.configs == 'all' else args.configs.split(',')
    del flush
    torch.cuda.empty_cache()
    for name in want:
      if world > 1 and name in ('c1', 'c4'):
        continue
      t0 = time.perf_counter()
      try:
        per_config[name] = getattr(bench_configs, 'run_' + name)(tfp, env, cpu=not args.no_cpu_baseline)
      except Exception as e:  # pylint: disable=broad-except
        per_config[name] = {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}
      log('[rank %d] %s: %.1fs %s' % (rank, name, time.perf_counter() - t0,
                                      {k: v for k, v in per_config[name].items() if k in ('value', 'error', 'min_ess_per_sec')}))
      torch.cuda.empty_cache()

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  pk, pk_src = peaks()
  # roofline of the dominant kernel (tile_nuts_async_kernel): ALGORITHMIC flops = 2*D^2 per gradient
  # evaluation, against the dense TF32 tensor peak (= 1/2 of the measured sustained bf16 peak), DESIGN.md section 5.
  flops = 2.0 * D * D * n_grad
  avg_launch_s = total_s
  achieved = flops / avg_launch_s / 1e12
  peak = 0.5 * pk.get('bf16_tflops_sustained', pk.get('bf16_tflops'))
  roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
              'frac_of_3xtf32_peak': achieved / (peak / 3.0),   # FP32-accurate split: 3 tensor-core passes per flop
              'traffic': NCU_DRAM_BYTES_PER_TRANSITION * args.steps,
              'traffic_source': 'ncu --set full of this kernel (profiles/r02_tile_nuts_async_ncu_full_summary.csv): '
                                'dram read+write bytes per transition of 16,384 chains, scaled to the K of this launch',
              'peak_source': pk_src + ': 0.5 x bf16_tflops_sustained (dense TF32)',
              'kernel': 'tile_nuts_async_kernel (tcgen05 kind::tf32; 3xTF32 split x 112/100 padding x two half-rows per '
                        'chain: the tensor pipe executes ~7x the algorithmic flops; ncu: tensor pipe 20 % active; the '
                        'kernel is bound by the dependent latency of one leapfrog, DESIGN.md section 5)'}
  cpu_baseline = None
  if not args.no_cpu_baseline and world == 1:
    cpu_baseline, _, _ = cpu_reference_arm(3, 1, budget_s=20.0)
  line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
          'warmup': max(args.warmup, 3), 'ms_per_step': 1e3 * total_s / args.steps, 'higher_is_better': True,
          'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
          'clocks': clocks, 'gpu_launches': int(launches),
          'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': B * D * 4,
                  'd2h_bytes_per_step': B * D * 4 + B * 4, 'steps': e2e_steps,
                  'api': 'tfp.mcmc.sample_chain(num_results=1) per step incl. bootstrap_results',
                  'one_call': {'value': one_call_value, 'unit': UNIT,
                               'note': 'H2D initial state + ONE sample_chain(num_results=%d) + D2H of all states' % e2e_steps}},
          'roofline': roofline, 'cpu_baseline': cpu_baseline, 'step_size': eps,
          'leapfrogs_per_transition': n_grad / (B * args.steps),
          'timed_reps_ms': [1e3 * r[0] for r in reps],
          'per_launch_run': {'value': fused_value, 'unit': UNIT,
                             'note': 'same K transitions as K one_step launches, L2 flushed between launches'},
          'min_ess': ess_info, 'per_config': per_config}
  if ess_info:
    line['min_ess_per_sec'] = ess_info['min_ess_per_sec_cross_chain']
  if shard_parity is not None:
    line['shard_parity'] = shard_parity
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
