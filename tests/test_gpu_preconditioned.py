"""GPU tests of diagonal preconditioning and the adaptation stack around it:
  * pb2_running_moments_update / RunningVariance vs NumPy;
  * PreconditionedHamiltonianMonteCarlo / PreconditionedNoUTurnSampler (kernels run on u = x / s) against the oracle,
    which restates the reference's own form -- momentum z / s, velocity s^2 m, kinetic energy 1/2 sum s^2 m^2, U-turn
    test on <rho, velocity> (experimental/mcmc/preconditioned_hmc.py, preconditioned_nuts.py:694-705,964-1030);
  * DiagonalMassMatrixAdaptation's running variance and momentum update (diagonal_mass_matrix_adaptation.py:200-290);
  * windowed_adaptive_nuts on the ill-conditioned 100-d Gaussian: the adapted mass matrix is the target's marginal
    variance and mixing improves by a large factor at equal gradient evaluations (windowed_sampling.py:322-347,603-783).
"""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu

from oracle import mcmc as omcmc  # noqa: E402
from oracle import rng as orng  # noqa: E402
from oracle import targets as otargets  # noqa: E402


@pytest.fixture(scope='module')
def tfp():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  import probability_b200 as tfp_
  return tfp_


def dev():
  return torch.device('cuda', 0)


def t(a):
  return torch.tensor(np.asarray(a), device=dev())


@pytest.mark.parametrize('D,batches', [(10, [64, 1, 300]), (2519, [7, 33]), (100, [4096, 4096])])
def test_running_moments_match_numpy(tfp, D, batches):
  rng = np.random.default_rng(D)
  rv = tfp.experimental.stats.RunningVariance.from_shape([(D,)], dev(), was_list=False)
  allx = []
  for n in batches:
    x = (3.0 + rng.standard_normal((n, D)) * rng.uniform(0.1, 5.0, D)).astype(np.float32)
    allx.append(x)
    rv = rv.update(t(x))
  ref = np.concatenate(allx).astype(np.float64)
  assert int(rv.num_samples) == ref.shape[0]
  np.testing.assert_allclose(rv.mean.cpu().numpy(), ref.mean(0), rtol=2e-5, atol=2e-5)
  np.testing.assert_allclose(rv.variance().cpu().numpy(), ref.var(0), rtol=5e-4)
  # from_stats continues an existing estimate (sample_stats.py:182-206)
  rv2 = tfp.experimental.stats.RunningVariance.from_stats(ref.shape[0], t(ref.mean(0).astype(np.float32)),
                                                          t(ref.var(0).astype(np.float32)))
  x = rng.standard_normal((50, D)).astype(np.float32)
  ref2 = np.concatenate([ref, x])
  np.testing.assert_allclose(rv2.update(t(x)).variance().cpu().numpy(), ref2.var(0), rtol=1e-3)


def _targets(tfp, which):
  rng = np.random.default_rng(3)
  if which == 'eight_schools':
    tg, og = tfp.targets.EightSchools(), otargets.EightSchools()
    x = (np.array([0, 0] + [1] * 8) + 0.5 * rng.standard_normal((96, 10))).astype(np.float32)
  elif which == 'dense3':
    cov = np.array([[1.0, 0.5, 0.1], [0.5, 2.0, -0.3], [0.1, -0.3, 0.5]])
    loc = np.array([1.0, -2.0, 0.5])
    tg = tfp.targets.DenseGaussian(covariance=cov, loc=loc)
    og = otargets.DenseGaussian(tg.precision, tg.log_normalizer, loc)
    x = rng.standard_normal((96, 3)).astype(np.float32)
  elif which == 'dense100':      # 512 chains: the tcgen05 tile kernels; the gym target's spectrum, axis-aligned
    ev = tfp.targets.IllConditionedGaussian().covariance_eigenvalues
    cov = np.diag(ev[rng.permutation(100)])
    cov[0, 1] = cov[1, 0] = 0.3 * np.sqrt(cov[0, 0] * cov[1, 1])
    tg = tfp.targets.DenseGaussian(covariance=cov)
    tg.covariance = cov
    og = otargets.DenseGaussian(tg.precision, tg.log_normalizer)
    x = (rng.standard_normal((512, 100)) @ np.linalg.cholesky(cov).T).astype(np.float32)
  elif which == 'logistic5':
    X, y = tfp.targets.synthetic_logistic_data(37, 4, seed=1)
    tg = tfp.targets.LogisticRegression(X, y)
    og = otargets.LogisticRegression(tg.features_with_bias, y)
    x = (0.3 * rng.standard_normal((96, 5))).astype(np.float32)
  elif which == 'sv60':
    yv = tfp.targets.synthetic_sv_returns(T=60, seed=2)
    tg, og = tfp.targets.StochasticVolatility(yv), otargets.StochasticVolatility(yv)
    x = (0.3 * rng.standard_normal((33, 63))).astype(np.float32)
    x[:, 0] += 2.0
    x[:, 1] += 5.0
  else:
    raise KeyError(which)
  var = rng.uniform(0.2, 4.0, x.shape[1]).astype(np.float32)
  if which == 'dense100':
    var = np.diag(tg.covariance).astype(np.float32)       # what adaptation converges to
  return tg, og, x, var


def _parts(tg, a):
  sizes = tg.part_sizes
  out, off = [], 0
  for n in sizes:
    piece = a[:, off:off + n]
    out.append(t(piece[:, 0] if n == 1 and len(sizes) > 1 else piece))
    off += n
  return out if len(out) > 1 else out[0]


def _var_parts(tg, var):
  sizes = tg.part_sizes
  out, off = [], 0
  for n in sizes:
    v = var[off:off + n]
    out.append(t(v[0] if n == 1 and len(sizes) > 1 else v))
    off += n
  return out


def _flat(state):
  if isinstance(state, (list, tuple)):
    return torch.cat([s.reshape(s.shape[0], -1) for s in state], 1).cpu().numpy()
  return state.cpu().numpy()


@pytest.mark.parametrize('which,eps,L', [('eight_schools', 0.3, 3), ('dense3', 0.5, 4), ('logistic5', 0.15, 3),
                                        ('sv60', 0.02, 3), ('dense100', 0.3, 5)])
def test_preconditioned_hmc_matches_oracle(tfp, which, eps, L):
  tg, og, x, var = _targets(tfp, which)
  md = tfp.experimental.mcmc.DiagonalMomentum(_var_parts(tg, var))
  k = tfp.experimental.mcmc.PreconditionedHamiltonianMonteCarlo(tg, step_size=eps, num_leapfrog_steps=L,
                                                                momentum_distribution=md)
  state = _parts(tg, x)
  seed = orng.key(13)
  new_state, kr = k.one_step(state, k.bootstrap_results(state), seed=seed)
  lp0, g0 = og.logp_grad(x)
  ref = omcmc.hmc_one_step(og, x, lp0, g0, eps, L, seed, inv_mass=var)
  np.testing.assert_allclose(_flat(kr.proposed_results.initial_momentum), ref['initial_momentum'], rtol=5e-6, atol=1e-6)
  np.testing.assert_allclose(_flat(kr.proposed_state), ref['proposed_state'], rtol=2e-4, atol=2e-4)
  np.testing.assert_allclose(kr.log_accept_ratio.cpu().numpy(), ref['log_accept_ratio'], rtol=2e-3, atol=3e-3)
  agree = kr.is_accepted.cpu().numpy() == ref['is_accepted']
  assert agree.mean() >= 0.97
  np.testing.assert_allclose(_flat(new_state)[agree], ref['state'][agree], rtol=2e-4, atol=2e-4)
  np.testing.assert_allclose(_flat(kr.proposed_results.final_momentum), ref['final_momentum'], rtol=2e-3, atol=2e-3)
  assert kr.accepted_results.momentum_distribution is md


@pytest.mark.parametrize('which,eps,depth', [('eight_schools', 0.25, 6), ('dense3', 0.4, 5), ('logistic5', 0.12, 5),
                                            ('sv60', 0.03, 4), ('dense100', 0.5, 8)])
def test_preconditioned_nuts_matches_oracle(tfp, which, eps, depth):
  tg, og, x, var = _targets(tfp, which)
  md = tfp.experimental.mcmc.DiagonalMomentum(_var_parts(tg, var))
  k = tfp.experimental.mcmc.PreconditionedNoUTurnSampler(tg, step_size=eps, max_tree_depth=depth,
                                                         momentum_distribution=md)
  state = _parts(tg, x)
  seed = orng.key(29)
  new_state, kr = k.one_step(state, k.bootstrap_results(state), seed=seed)
  lp0, g0 = og.logp_grad(x)
  ref = omcmc.nuts_one_step(og, x, lp0, g0, eps, seed, max_tree_depth=depth, inv_mass=var)
  nl = kr.leapfrogs_taken.cpu().numpy()
  same = nl == ref['leapfrogs_taken']
  print('preconditioned NUTS %s: identical trees %.4f (mean leapfrogs %.1f)' % (which, same.mean(), nl.mean()))
  assert same.mean() >= 0.97, (nl, ref['leapfrogs_taken'])
  close = np.isclose(_flat(new_state), ref['state'], rtol=3e-3, atol=3e-3).all(axis=1)
  assert close[same].mean() >= 0.97
  for f in ('is_accepted', 'reach_max_depth', 'has_divergence'):
    assert (getattr(kr, f).cpu().numpy() == ref[f])[same & close].all(), f
  np.testing.assert_allclose(kr.log_accept_ratio.cpu().numpy()[same & close], ref['log_accept_ratio'][same & close],
                             rtol=3e-3, atol=3e-3)
  assert kr.momentum_distribution is md
  if which == 'dense100':
    assert nl.mean() < 40      # preconditioned with the marginal variances the trees are short (identity: ~270)


def test_preconditioned_fused_equals_step_loop(tfp):
  tg, og, x, var = _targets(tfp, 'eight_schools')
  md = tfp.experimental.mcmc.DiagonalMomentum(_var_parts(tg, var))
  k = tfp.experimental.mcmc.PreconditionedNoUTurnSampler(tg, step_size=0.2, max_tree_depth=5, momentum_distribution=md)
  state = _parts(tg, x)
  a = tfp.mcmc.sample_chain(6, state, kernel=k, seed=4, trace_fn=lambda _, kr: kr.leapfrogs_taken)
  b = tfp.mcmc.sample_chain(6, state, kernel=k, seed=4, trace_fn=lambda _, kr: kr.leapfrogs_taken + 0)   # step loop
  np.testing.assert_array_equal(a.trace.cpu().numpy(), b.trace.cpu().numpy())
  # (the step loop crosses the ABI in the original coordinates every transition: x -> u = x / s -> x = s u costs an ulp
  # per step; the fused run stays in u)
  for u, v in zip(a.all_states, b.all_states):
    np.testing.assert_allclose(u.cpu().numpy(), v.cpu().numpy(), rtol=1e-3, atol=2e-5)


def test_diagonal_mass_matrix_adaptation(tfp):
  """The running variance is the variance of all visited states (chains x steps), and the momentum distribution is
  replaced when step == num_estimation_steps (diagonal_mass_matrix_adaptation.py:280-287)."""
  tg, og, x, _ = _targets(tfp, 'dense3')
  st = t(x)
  exp = tfp.experimental
  inner = exp.mcmc.PreconditionedHamiltonianMonteCarlo(tg, step_size=0.4, num_leapfrog_steps=3)
  rv0 = exp.stats.RunningVariance.from_shape([(3,)], dev(), was_list=False)
  k = exp.mcmc.DiagonalMassMatrixAdaptation(inner, initial_running_variance=rv0, num_estimation_steps=5)
  kr = k.bootstrap_results(st)
  seed = tfp.random.sanitize_seed(3)
  seen = []
  for i in range(8):
    s, seed = tfp.random.split_seed(seed)
    st, kr = k.one_step(st, kr, seed=s)
    if i < 5:
      seen.append(st.cpu().numpy())
    got = [v.cpu().numpy() for v in kr.inner_results.accepted_results.momentum_distribution.variance()][0]
    if i < 4:
      np.testing.assert_array_equal(got, np.ones(3, np.float32))        # not yet replaced
  ref_var = np.concatenate(seen).astype(np.float64).var(0)
  np.testing.assert_allclose(got, ref_var, rtol=1e-4)
  assert int(kr.running_variance.num_samples) == 5 * x.shape[0] and kr.step == 8


def test_windowed_adaptive_nuts_ill_conditioned_gaussian(tfp):
  """100-d Gaussian with the spectrum of C2's target (condition number 1.3e5), axis-aligned: after windowed adaptation
  the mass matrix is the marginal variance, split R-hat < 1.01 and the min-ESS per gradient evaluation is many times
  that of the identity mass matrix.  (C2's own target is RANDOMLY ROTATED: its correlation matrix has condition number
  1.6e5, so no diagonal mass matrix can help there -- in the reference either; DESIGN.md.)"""
  ev = tfp.targets.IllConditionedGaussian().covariance_eigenvalues
  tg = tfp.targets.DenseGaussian(covariance=np.diag(ev[np.random.default_rng(1).permutation(100)]))
  tg.covariance = np.diag(1.0 / np.diag(tg.precision.astype(np.float64)))
  B, n_draws = 512, 100
  rng = np.random.default_rng(0)
  x0 = (rng.standard_normal((B, 100)) @ np.linalg.cholesky(tg.covariance).T).astype(np.float32)
  draws, trace = tfp.experimental.mcmc.windowed_adaptive_nuts(n_draws, tg, current_state=t(x0),
                                                              num_adaptation_steps=500, seed=11)
  assert tuple(draws.shape) == (n_draws, B, 100)
  vs = trace['variance_scaling'][0].cpu().numpy()
  truth = np.diag(tg.covariance)
  assert np.max(np.abs(np.log(vs / truth))) < 0.5            # the adapted mass matrix ~ marginal variances
  rhat = tfp.mcmc.potential_scale_reduction(draws, split_chains=True).cpu().numpy()
  print('windowed NUTS: max split R-hat %.4f, step size %.3f, mean leapfrogs %.1f' % (
      rhat.max(), float(trace['step_size'][0]), trace['n_steps'].float().mean().item()))
  assert rhat.max() < 1.01
  ess_w = tfp.mcmc.effective_sample_size(draws, cross_chain_dims=1, filter_beyond_positive_pairs=True,
                                         filter_threshold=None).cpu().numpy().min()
  grads_w = trace['n_steps'].float().sum().item()
  # identity mass matrix at its own adapted step size, same number of draws
  nuts = tfp.mcmc.NoUTurnSampler(tg, step_size=0.158, max_tree_depth=10)
  da = tfp.mcmc.DualAveragingStepSizeAdaptation(nuts, num_adaptation_steps=150)
  r = tfp.mcmc.sample_chain(n_draws, t(x0), kernel=da, num_burnin_steps=160, seed=12,
                            trace_fn=lambda _, kr: kr.inner_results.leapfrogs_taken)
  ess_i = tfp.mcmc.effective_sample_size(r.all_states, cross_chain_dims=1, filter_beyond_positive_pairs=True,
                                         filter_threshold=None).cpu().numpy().min()
  grads_i = r.trace.float().sum().item()
  gain = (ess_w / grads_w) / (ess_i / grads_i)
  print('   min-ESS per gradient: windowed %.3e vs identity %.3e (x%.1f)' % (ess_w / grads_w, ess_i / grads_i, gain))
  assert gain > 5.0
  # posterior variances within 4 MCSE-ish of the truth
  var_hat = draws.reshape(-1, 100).var(0).cpu().numpy()
  np.testing.assert_allclose(var_hat, truth, rtol=0.15)


def test_sample_fold_streaming_reducers(tfp):
  """experimental/mcmc/sample_fold.py + reducers: mean / variance / R-hat of a run computed from running moments, chunk
  by chunk, equal the statistics of the materialised history of the same chunked run."""
  tg, og, x, _ = _targets(tfp, 'eight_schools')
  exp = tfp.experimental.mcmc
  state = _parts(tg, x)
  k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.3, num_leapfrog_steps=3)
  n_steps = 57
  red = [exp.ExpectationsReducer(), exp.VarianceReducer(), exp.PotentialScaleReductionReducer()]
  chunk_bytes = 10 * x.shape[0] * x.shape[1] * 4       # chunks of 10 steps: 6 chunks, the last one short
  out = exp.sample_fold(n_steps, state, kernel=k, reducer=red, seed=21, experimental_chunk_bytes=chunk_bytes)
  mean, var, rhat = out.reduction_results
  # the same chunked run, materialised
  seed = tfp.random.sanitize_seed(21, salt='mcmc.sample_fold')
  st, pkr, hist, done = state, k.bootstrap_results(state), [], 0
  while done < n_steps:
    n = min(10, n_steps - done)
    cs, seed = tfp.random.split_seed(seed)
    r = tfp.mcmc.sample_chain(n, st, previous_kernel_results=pkr, kernel=k, trace_fn=None, seed=cs,
                              return_final_kernel_results=True)
    hist.append(_flat([s.reshape(-1, *s.shape[2:]) for s in r.all_states]).reshape(n, x.shape[0], -1))
    st, pkr, done = [s[-1] for s in r.all_states], r.final_kernel_results, done + n
  h = np.concatenate(hist).astype(np.float64)            # [n_steps, B, D]
  np.testing.assert_allclose(_flat(mean), h.mean(0), rtol=1e-4, atol=1e-4)
  np.testing.assert_allclose(_flat(var), h.var(0), rtol=2e-3, atol=1e-5)
  ref_rhat = tfp.mcmc.potential_scale_reduction(t(h.astype(np.float32)), split_chains=False).cpu().numpy()
  got_rhat = np.concatenate([np.atleast_1d(v.cpu().numpy()).reshape(-1) for v in rhat])
  np.testing.assert_allclose(got_rhat, ref_rhat, rtol=2e-3)
  for a, b in zip(out.end_state, st):
    np.testing.assert_array_equal(a.cpu().numpy(), b.cpu().numpy())


def test_covariance_reducer_with_reductions_and_step_kernel(tfp):
  """covariance_reducer.py / with_reductions.py:41 / step.py:29: (i) the per-chain running covariance of a chunked
  `sample_fold` run equals np.cov of the materialised history; (ii) a `WithReductions` kernel driven by `step_kernel`
  (one transition + one fold per step) ends in the same state as the step loop and its reducers hold the statistics of
  exactly the states it visited; (iii) `step_kernel` follows the reference's seed loop."""
  exp = tfp.experimental.mcmc
  tg = tfp.targets.IllConditionedGaussian()
  rng = np.random.default_rng(4)
  x0 = t((rng.standard_normal((24, 100)) @ np.linalg.cholesky(tg.covariance).T).astype(np.float32))
  k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.5, num_leapfrog_steps=4)
  n_steps = 33
  chunk_bytes = 8 * 24 * 100 * 4
  out = exp.sample_fold(n_steps, x0, kernel=k, reducer=[exp.CovarianceReducer(ddof=1), exp.VarianceReducer(ddof=1)],
                        seed=3, experimental_chunk_bytes=chunk_bytes)
  cov, var = out.reduction_results
  seed = tfp.random.sanitize_seed(3, salt='mcmc.sample_fold')
  st, pkr, hist, done = x0, k.bootstrap_results(x0), [], 0
  while done < n_steps:
    n = min(8, n_steps - done)
    cs, seed = tfp.random.split_seed(seed)
    r = tfp.mcmc.sample_chain(n, st, previous_kernel_results=pkr, kernel=k, trace_fn=None, seed=cs,
                              return_final_kernel_results=True)
    hist.append(r.all_states.cpu().numpy())
    st, pkr, done = r.all_states[-1], r.final_kernel_results, done + n
  h = np.concatenate(hist).astype(np.float64)                  # [n_steps, B, D]
  assert tuple(cov.shape) == (24, 100, 100)
  for b in (0, 7, 23):
    ref = np.cov(h[:, b, :], rowvar=False, ddof=1)
    np.testing.assert_allclose(cov[b].cpu().numpy(), ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max())
  np.testing.assert_allclose(torch.diagonal(cov, dim1=1, dim2=2).cpu().numpy(), var.cpu().numpy(), rtol=1e-3, atol=1e-5)
  with pytest.raises(NotImplementedError):
    exp.CovarianceReducer(event_ndims=None)
  # (ii) WithReductions + step_kernel == the step loop with the same seeds
  reds = [exp.ExpectationsReducer(), exp.CovarianceReducer()]
  wk = exp.WithReductions(k, reds)
  assert wk.is_calibrated and wk.inner_kernel is k
  end, kr = exp.step_kernel(9, x0, kernel=wk, return_final_kernel_results=True, seed=11)
  seed = tfp.random.sanitize_seed(11, salt='mcmc_step_kernel')
  st, pkr, visited = x0, k.bootstrap_results(x0), []
  for _ in range(9):
    ss, seed = tfp.random.split_seed(seed)
    st, pkr = k.one_step(st, pkr, seed=ss)
    visited.append(st.cpu().numpy())
  np.testing.assert_array_equal(end.cpu().numpy(), st.cpu().numpy())
  v = np.stack(visited).astype(np.float64)
  mean_w, cov_w = [r.finalize(s) for r, s in zip(reds, kr.reduction_results)]
  np.testing.assert_allclose(mean_w.cpu().numpy(), v.mean(0), rtol=1e-4, atol=1e-4)
  ref = np.cov(v[:, 5, :], rowvar=False, ddof=0)
  np.testing.assert_allclose(cov_w[5].cpu().numpy(), ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max())
  # the bare driver returns only the state
  np.testing.assert_array_equal(exp.step_kernel(9, x0, kernel=k, seed=11).cpu().numpy(), st.cpu().numpy())


def test_sample_discarding_kernel_is_burnin_and_thinning(tfp):
  """sample_discarding_kernel.py:40-175: the first step takes burn-in + thinning + 1 inner transitions, later steps
  thinning + 1; wrapped in WithReductions the reducers see the surviving states only."""
  exp = tfp.experimental.mcmc
  tg = tfp.targets.EightSchools()
  x0 = t((np.array([0, 0] + [1] * 8) + 0.2 * np.random.default_rng(1).standard_normal((32, 10))).astype(np.float32))
  k = tfp.mcmc.HamiltonianMonteCarlo(tg, step_size=0.3, num_leapfrog_steps=3)
  dk = exp.SampleDiscardingKernel(k, num_burnin_steps=4, num_steps_between_results=2)
  kr = dk.bootstrap_results(x0)
  assert kr.call_counter == 0 and dk.is_calibrated
  s1, kr1 = dk.one_step(x0, kr, seed=(5, 6))
  s2, kr2 = dk.one_step(s1, kr1, seed=(7, 8))
  assert kr2.call_counter == 2
  # the same transitions by hand: 4 + 2 + 1 steps seeded from (5, 6), then 2 + 1 from (7, 8)
  a = exp.step_kernel(7, x0, kernel=k, seed=(5, 6))
  np.testing.assert_array_equal(s1.cpu().numpy(), a.cpu().numpy())
  b = exp.step_kernel(3, a, kernel=k, seed=(7, 8))     # kernel results after a bootstrap at `a` equal the carried ones
  np.testing.assert_array_equal(s2.cpu().numpy(), b.cpu().numpy())
  # the onion: reducers over the surviving states
  wk = exp.WithReductions(dk, exp.ExpectationsReducer())
  end, wkr = exp.step_kernel(5, x0, kernel=wk, return_final_kernel_results=True, seed=3)
  assert wkr.inner_results.call_counter == 5
  assert float(wkr.reduction_results.count) == 5.0
