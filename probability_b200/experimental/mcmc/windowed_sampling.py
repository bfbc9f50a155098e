"""Windowed adaptation: `windowed_adaptive_nuts` / `windowed_adaptive_hmc` (tfp/experimental/mcmc/windowed_sampling.py:
603-960; schedule :322-347; WindowedAdaptation.one_step :418-506).

Schedule (for the default 500 adaptation steps: 75 fast / 25, 50, 100, 200 slow / 50 fast): the step size is adapted by
dual averaging in every window, the diagonal mass matrix is re-estimated from the draws of each slow window (all
chains x all steps of the window are the observations of a fresh running variance) and takes effect when the window
ends; dual averaging restarts from the current step size at the start of the 2nd, 3rd and 4th slow window and of the
last window -- exactly the resets of WindowedAdaptation.one_step.

Execution: every window is ONE fused `sample_chain` call (pb2_run: persistent transition kernels + device-side dual
averaging), the window's draws are folded into the running variance by pb2_running_moments_update, and the mass matrix
enters the next window as pb2_run_cfg.d_momentum_scale.  Deviations from the reference: the model is a
probability_b200 Target in unconstrained coordinates instead of a JointDistribution (no bijector / pinning step), and
every window derives its seed from `seed` by fold_in(window index) (the reference threads one seed through one
sample_chain over all steps).
"""
import collections

import numpy as np

from probability_b200 import random as pb_random
from probability_b200.experimental.mcmc import preconditioning
from probability_b200.experimental.stats import RunningVariance
from probability_b200.mcmc import _engine
from probability_b200.mcmc import dual_averaging_step_size_adaptation as dassa
from probability_b200.mcmc import sample as sample_lib

WindowedSamples = collections.namedtuple('WindowedSamples', ['all_states', 'trace', 'final_kernel_results'])


def _get_window_sizes(num_adaptation_steps):
  """windowed_sampling.py:322-347: first (fast) window, initial slow window, last (fast) window."""
  slow_window_size = num_adaptation_steps // 20
  first_window_size = 3 * slow_window_size
  last_window_size = num_adaptation_steps - 15 * slow_window_size - first_window_size
  return first_window_size, slow_window_size, last_window_size


def _get_step_size(num_dims):
  """:566-589: 0.5 * (number of event dimensions) ** -0.25."""
  return 0.5 * float(num_dims) ** -0.25


def default_nuts_trace_fn(state, is_tuning, pkr):
  """:46-57 (the reference's `bijector` argument has no counterpart: states are in the target's own coordinates)."""
  del state
  return {'step_size': pkr.step_size, 'n_steps': pkr.leapfrogs_taken, 'tune': is_tuning,
          'target_log_prob': pkr.target_log_prob, 'diverging': pkr.has_divergence,
          'accept_ratio': pkr.log_accept_ratio.clamp(max=0.).exp(),
          'variance_scaling': pkr.momentum_distribution.variance(), 'is_accepted': pkr.is_accepted}


def default_hmc_trace_fn(state, is_tuning, pkr):
  """:60-71."""
  del state
  return {'step_size': pkr.accepted_results.step_size, 'tune': is_tuning,
          'target_log_prob': pkr.accepted_results.target_log_prob,
          'diverging': ~pkr.log_accept_ratio.isfinite() | (pkr.log_accept_ratio < -1000.),
          'accept_ratio': pkr.log_accept_ratio.clamp(max=0.).exp(),
          'variance_scaling': pkr.accepted_results.momentum_distribution.variance(), 'is_accepted': pkr.is_accepted}


def _windowed(kind, n_draws, target, n_chains, num_adaptation_steps, current_state, init_step_size,
              dual_averaging_kwargs, kernel_kwargs, seed, return_final_kernel_results, discard_tuning):
  import torch
  seed = pb_random.sanitize_seed(seed, salt='windowed_adaptive_' + kind)
  D = target.dim
  # chains sharded over ranks (experimental_chain_shard): step size and mass matrix are adapted on the statistics of ALL
  # ranks' chains -- dual averaging reduces its accept statistic over the ranks, the running variance of every window is
  # merged over the ranks in rank order -- like the reference's `experimental_chain_axis_names` (windowed_sampling.py)
  shard = kernel_kwargs.get('experimental_chain_shard')
  world = dassa.world_of('ranks' if shard is not None else None)
  if current_state is None:
    # init_near_unconstrained_zero: Uniform(-2, 2) in the unconstrained space (drawn for the global batch: a shard
    # starts from its rows of it)
    n_all = int(shard.num_chains_global) if shard is not None else n_chains
    u = pb_random.uniform((n_all, D), seed=pb_random.fold_in(seed, 1000))
    if shard is not None:
      u = u[int(shard.chain_offset):int(shard.chain_offset) + n_chains]
    x = (4.0 * u - 2.0).contiguous()
    shapes, was_list = [(D,)], False
    state = x
  else:
    x, shapes, was_list = _engine.flatten_state(current_state)
    state = current_state
    n_chains = x.shape[0]
  dev = x.device
  step_size = float(init_step_size) if init_step_size is not None else _get_step_size(D)
  da_kw = dict(target_accept_prob=0.85)
  da_kw.update(dual_averaging_kwargs or {})
  da_kw.pop('num_adaptation_steps', None)
  if world is not None:
    da_kw.setdefault('experimental_reduce_chain_axis_names', 'ranks')
  first, slow, last = _get_window_sizes(int(num_adaptation_steps))
  md = preconditioning.DiagonalMomentum([torch.ones(s if len(s) else (), device=dev) for s in shapes])

  def make_kernel(eps, momentum):
    if kind == 'nuts':
      return preconditioning.PreconditionedNoUTurnSampler(target, eps, momentum_distribution=momentum, **kernel_kwargs)
    return preconditioning.PreconditionedHamiltonianMonteCarlo(target, eps, momentum_distribution=momentum,
                                                               **kernel_kwargs)

  def inner_of(kr):
    return kr.inner_results

  tune_states, tune_lens = [], []

  def run_window(idx, length, n_var):
    """`length` transitions with dual averaging over all of them; the last `n_var` draws re-estimate the variance."""
    nonlocal state, step_size, md
    if length <= 0:
      return
    k = dassa.DualAveragingStepSizeAdaptation(make_kernel(step_size, md), num_adaptation_steps=length, **da_kw)
    res = sample_lib.sample_chain(length, state, kernel=k, trace_fn=None, return_final_kernel_results=True,
                                  seed=pb_random.fold_in(seed, idx))
    st = res.all_states
    state = [s[-1] for s in st] if was_list else st[-1]
    step_size = float(dassa._flat(res.final_kernel_results.new_step_size)[0])
    if n_var > 0:
      rv = RunningVariance.from_shape(shapes, dev, was_list)
      tail = [s[length - n_var:] for s in st] if was_list else st[length - n_var:]
      rv = rv.update(tail)
      if world is not None:
        rv = rv.merged_over_ranks(world)
      md = preconditioning.DiagonalMomentum(rv.variance() if was_list else [rv.variance()])
    if not discard_tuning:
      tune_states.append(st)
      tune_lens.append(length)

  if slow > 0:
    run_window(0, first + slow, slow)          # fast window + first slow window share one dual-averaging run (:425-441)
    run_window(1, 2 * slow, 2 * slow)
    run_window(2, 4 * slow, 4 * slow)
    run_window(3, 8 * slow, 8 * slow)
    run_window(4, last, 0)
  else:
    run_window(0, int(num_adaptation_steps), 0)

  kernel = make_kernel(step_size, md)
  trace_fn = default_nuts_trace_fn if kind == 'nuts' else default_hmc_trace_fn
  res = sample_lib.sample_chain(int(n_draws), state, kernel=kernel, return_final_kernel_results=True,
                                trace_fn=(lambda s, kr: kr), seed=pb_random.fold_in(seed, 5))
  kr = res.trace
  if kind == 'nuts':
    trace = {'step_size': torch.full((int(n_draws),), step_size, device=dev), 'n_steps': kr.leapfrogs_taken,
             'tune': torch.zeros(int(n_draws), dtype=torch.bool, device=dev), 'target_log_prob': kr.target_log_prob,
             'diverging': kr.has_divergence, 'accept_ratio': kr.log_accept_ratio.clamp(max=0.).exp(),
             'variance_scaling': md.variance(), 'is_accepted': kr.is_accepted}
  else:
    trace = {'step_size': torch.full((int(n_draws),), step_size, device=dev),
             'tune': torch.zeros(int(n_draws), dtype=torch.bool, device=dev),
             'target_log_prob': kr.accepted_results.target_log_prob,
             'diverging': ~kr.log_accept_ratio.isfinite() | (kr.log_accept_ratio < -1000.),
             'accept_ratio': kr.log_accept_ratio.clamp(max=0.).exp(), 'variance_scaling': md.variance(),
             'is_accepted': kr.is_accepted}
  draws = res.all_states
  if not discard_tuning and tune_states:
    if was_list:
      draws = [torch.cat([t[i] for t in tune_states] + [draws[i]]) for i in range(len(draws))]
    else:
      draws = torch.cat(tune_states + [draws])
    trace['num_tuning_steps'] = int(sum(tune_lens))
  if return_final_kernel_results:
    return WindowedSamples(draws, trace, res.final_kernel_results)
  return draws, trace


def windowed_adaptive_nuts(n_draws, target, *, n_chains=64, num_adaptation_steps=500, current_state=None,
                           init_step_size=None, dual_averaging_kwargs=None, max_tree_depth=10, max_energy_diff=500.,
                           unrolled_leapfrog_steps=1, parallel_iterations=10, return_final_kernel_results=False,
                           discard_tuning=True, experimental_chain_shard=None, seed=None):
  """Adapt and sample with NUTS (windowed_sampling.py:603-783).  Returns `(draws, trace)`; `trace` is the dict of
  `default_nuts_trace_fn` over the draws (step_size, n_steps, tune, target_log_prob, diverging, accept_ratio,
  variance_scaling, is_accepted)."""
  kw = dict(max_tree_depth=max_tree_depth, max_energy_diff=max_energy_diff,
            unrolled_leapfrog_steps=unrolled_leapfrog_steps, parallel_iterations=parallel_iterations,
            experimental_chain_shard=experimental_chain_shard)
  return _windowed('nuts', n_draws, target, n_chains, num_adaptation_steps, current_state, init_step_size,
                   dual_averaging_kwargs, kw, seed, return_final_kernel_results, discard_tuning)


def windowed_adaptive_hmc(n_draws, target, *, num_leapfrog_steps, n_chains=64, num_adaptation_steps=500,
                          current_state=None, init_step_size=None, dual_averaging_kwargs=None,
                          return_final_kernel_results=False, discard_tuning=True, experimental_chain_shard=None,
                          seed=None):
  """Adapt and sample with HMC (windowed_sampling.py:786-960)."""
  kw = dict(num_leapfrog_steps=num_leapfrog_steps, experimental_chain_shard=experimental_chain_shard)
  return _windowed('hmc', n_draws, target, n_chains, num_adaptation_steps, current_state, init_step_size,
                   dual_averaging_kwargs, kw, seed, return_final_kernel_results, discard_tuning)
