"""CPU tests that PIN the oracle against the reference's own known-answer tests and
invariants -- the oracle is only trusted as a checker after these."""
import numpy as np
import pytest

from oracle import diagnostic as odiag
from oracle import mcmc as omcmc
from oracle import rng as orng
from oracle import targets as otargets


# ---- RNG: Random123 KATs, jax-documented values, and the JAX outputs held by the reference's notebook -------------------------------------
def test_threefry_random123_kats():
  cases = [((0, 0), (0, 0), (0x6b200159, 0x99ba4efe)),
           ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
           ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))]
  for k, c, out in cases:
    o0, o1 = orng.threefry2x32(k[0], k[1], c[0], c[1])
    assert (int(o0), int(o1)) == out


def test_jax_documented_values():
  np.testing.assert_array_equal(orng.split(orng.key(0), 2, orng.ORIGINAL),
                                [[4146024105, 967050713], [2718843009, 1272950319]])
  np.testing.assert_array_equal(orng.split(orng.key(0), 2, orng.PARTITIONABLE),
                                [[1797259609, 2579123966], [928981903, 3453687069]])
  np.testing.assert_allclose(orng.normal(orng.key(0), (1,), orng.ORIGINAL), [-0.20584227], rtol=1e-6)


def test_rng_reproduces_the_jax_outputs_the_reference_keeps():
  """tests/golden/jax_notebook_rng.json: outputs of a live JAX run held by the reference's own executed notebook
  (examples/jupyter_notebooks/TensorFlow_Probability_on_JAX.ipynb cells 35, 49, 96-102; extracted by make_golden.py).
  They pin the threefry stream, the key split and the bits -> uniform -> normal transform of the seed contract on
  reference-held vectors: PRNGKey(0), split, normal on the key and on both children, and tfd.Normal.sample(seed=key)
  == jax.random.normal(key) (the JAX substrate's samplers.normal adds no salt, internal/samplers.py:308-325)."""
  import json, os
  g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'jax_notebook_rng.json')))
  k = orng.key(0)
  np.testing.assert_array_equal(k, g['prng_key_0'])
  np.testing.assert_allclose(orng.normal(k, (1,), orng.ORIGINAL)[0], g['normal_key_0'], rtol=2e-7)
  np.testing.assert_allclose(orng.normal(k, (1,), orng.ORIGINAL)[0], g['tfd_normal_sample_key_0'], rtol=2e-7)
  kids = orng.split(k, 2, orng.ORIGINAL)
  np.testing.assert_array_equal(kids, g['split_key_0'])
  z = [orng.normal(kk, (1,), orng.ORIGINAL)[0] for kk in kids]
  np.testing.assert_allclose(z, g['normal_split_keys'], rtol=2e-7)
  loc, var = g['split_normal_exp_mean_variance']        # Normal(normal(k1), exp(normal(k2))): mean, variance
  np.testing.assert_allclose(z[0], loc, rtol=2e-7)
  np.testing.assert_allclose(np.exp(np.float32(z[1])) ** 2, var, rtol=1e-6)
  # multi-element streams and a 2-d shape (the counter layout of the original generator), from two more notebooks
  m = g['more']
  np.testing.assert_allclose(orng.uniform(k, (2,), layout=orng.ORIGINAL), m['uniform_1x2_key_0']['value'], rtol=2e-7)
  np.testing.assert_allclose(np.exp(orng.normal(k, (10,), orng.ORIGINAL)).reshape(2, 5),
                             m['exp_normal_2x5_key_0']['value'], rtol=5e-7)
  # a TPU run: same bits, that platform's erfinv (last digits differ; see the note in the fixture)
  np.testing.assert_allclose(orng.normal(k, (8,), orng.ORIGINAL), m['normal_8_key_0_tpu']['value'], rtol=2e-5)


def test_hmc_run_held_by_the_reference_notebook():
  """tests/golden/tf_notebook_hmc.json: the executed cell of TFP_Release_Notebook_0_11_0.ipynb --
  sample_chain(5, zeros([3]), HamiltonianMonteCarlo(lambda x: -(x - .2)**2, step_size=1., num_leapfrog_steps=2),
  num_burnin_steps=100).  Two unit leapfrogs map x -> 0.4 - x, m -> -m for ANY momentum, so the run does not depend on
  the generator: the oracle's leapfrog (leapfrog_integrator.py:222-316), accept step (metropolis_hastings.py:160-254)
  and burn-in indexing (sample.py:311-383) reproduce the reference's printed states."""
  import json, os
  g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'tf_notebook_hmc.json')))

  class Quadratic(object):
    dim, part_sizes = 1, [1]

    def logp_grad(self, x):
      x = np.asarray(x, np.float32)
      return (-(x[:, 0] - np.float32(0.2)) ** 2).astype(np.float32), (np.float32(-2.0) * (x - np.float32(0.2)))

  states, trace, _ = omcmc.sample_chain(Quadratic(), 'hmc', np.zeros((3, 1), np.float32), g['num_results'],
                                        num_burnin_steps=g['num_burnin_steps'], step_size=g['step_size'],
                                        num_leapfrog_steps=g['num_leapfrog_steps'], seed=(1, 2))
  np.testing.assert_allclose(states[:, :, 0], g['all_states'], atol=1e-5)
  assert all(t['is_accepted'].all() for t in trace)
  np.testing.assert_allclose(trace[0]['target_log_prob'], g['target_log_prob_first_result'], atol=2e-6)
  assert max(np.abs(t['log_acceptance_correction']).max() for t in trace) < 2e-6
  assert np.abs(np.asarray(g['log_acceptance_correction'])).max() < 2e-6


def test_sample_chain_salt():
  # samplers.py:159-163 salt for 'mcmc.sample_chain' and the derived first step seed
  assert orng.salt_int('mcmc.sample_chain') == 1365385517
  k = orng.sanitize_seed(17, salt='mcmc.sample_chain')
  np.testing.assert_array_equal(k, [2648418349, 421598061])
  np.testing.assert_array_equal(orng.split(k, 2, orng.ORIGINAL),
                                [[2659485321, 3880704392], [861450069, 1369571713]])


def test_uniform_and_normal_ranges():
  for layout in (0, 1):
    u = orng.uniform(orng.key(1), (100001,), layout=layout)
    assert u.min() >= 0 and u.max() < 1
    assert abs(u.mean() - 0.5) < 5e-3
    z = orng.normal(orng.key(2), (100001,), layout)
    assert np.isfinite(z).all()
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    r = orng.randint_bit(orng.key(3), 10001, layout)
    assert set(np.unique(r)) == {0, 1}
    np.testing.assert_array_equal(r, orng.randint(orng.key(3), 10001, 0, 2, layout))


# ---- NUTS instruction tables (nuts_test.py:191-215) -----------------------------------
def test_nuts_instruction_pins_depth4():
  write, read = omcmc.write_read_instructions(4)
  np.testing.assert_array_equal(write, [0, 4, 1, 4, 1, 4, 2, 4, 1, 4, 2, 4, 2, 4, 3, 4])
  np.testing.assert_array_equal(
      read, [[0, 0], [0, 1], [0, 0], [0, 2], [0, 0], [1, 2], [0, 0], [0, 3], [0, 0], [1, 2], [0, 0], [1, 3],
             [0, 0], [2, 3], [0, 0], [0, 4]])


@pytest.mark.parametrize('depth', range(1, 11))
def test_nuts_closed_form_equals_construction(depth):
  a = omcmc.write_read_instructions(depth)
  b = omcmc.write_read_closed_form(depth)
  np.testing.assert_array_equal(a[0], b[0])
  np.testing.assert_array_equal(a[1], b[1])


def test_product_tables_match(lib_built):
  from probability_b200.mcmc import nuts
  for depth in (1, 4, 10):
    w, r = nuts.generate_efficient_write_read_instruction(depth)
    a = omcmc.write_read_instructions(depth)
    np.testing.assert_array_equal(w, a[0])
    np.testing.assert_array_equal(r, a[1])
  np.testing.assert_array_equal(nuts.build_tree_uturn_instruction(3), omcmc.build_tree_uturn_instruction(3))


# ---- dual averaging (dual_averaging_step_size_adaptation_test.py:51-57) ----------------
@pytest.mark.parametrize('p,expect', [(0.70, 9.131008), (0.73, 9.642897), (0.74, 9.819825), (0.75, 10.0),
                                      (0.76, 10.183481)])
def test_dual_averaging_update_pins(p, expect):
  da = omcmc.DualAveraging(1.0, num_adaptation_steps=100)
  got = da.update(np.log(np.full(3, p, np.float32)))
  np.testing.assert_allclose(got, expect, rtol=2e-6)


def test_dual_averaging_list_step_pin():
  # testListStep (:177-200): per-part step sizes 0.1/0.2/0.3 with accept probs .74/.76/.76
  for s0, p, err in [(0.1, 0.74, 0.01), (0.2, 0.76, -0.01), (0.3, 0.76, -0.01)]:
    da = omcmc.DualAveraging(s0, num_adaptation_steps=1)
    da.update(np.log(np.float32([p])))
    # step == num_adaptation_steps: averaged step = exp(eta*log_step), eta = 1 at t = 1
    expect = np.exp(np.log(10. * s0) - err / ((10. + 1.) * 0.05))
    np.testing.assert_allclose(da.step_size, expect, rtol=1e-5)
    before = da.step_size
    da.update(np.log(np.float32([p])))
    assert da.step_size == before      # frozen afterwards


def test_dual_averaging_nonfinite_accept_is_zero_prob():
  da = omcmc.DualAveraging(1.0, 10)
  da.update(np.float32([np.nan, -np.inf, np.inf, 0.0]))   # probs: 0, 0, 0 (inf -> -inf), 1
  np.testing.assert_allclose(da.error_sum, 0.75 - 0.25, rtol=1e-6)


# ---- diagnostics (diagnostic_test.py:405-437, 653-703, 56-176) -------------------------
def test_rhat_pin():
  state = np.arange(15.).reshape(5, 3)
  n, m = 5., 3.
  b_div_n = np.var(state.mean(0), ddof=1)
  w = np.mean(np.var(state, axis=0, ddof=1))
  expect = ((m + 1) / m) * ((n - 1) / n * w + b_div_n) / w - (n - 1) / (m * n)
  np.testing.assert_allclose(odiag.potential_scale_reduction(state), expect, rtol=1e-12)
  assert odiag.potential_scale_reduction(np.arange(15).reshape(5, 3).astype(np.int64)).dtype == np.float64


@pytest.mark.parametrize('shape,axis', [((5,), None), ((5, 3), 0), ((5, 3), 1), ((4, 3, 2), (0, 1)),
                                        ((4, 3, 2), None), ((4, 3, 2), 2)])
def test_reduce_variance_vs_numpy(shape, axis):
  x = np.random.default_rng(0).standard_normal(shape)
  np.testing.assert_allclose(odiag.reduce_variance(x, axis=axis, biased=True), np.var(x, axis=axis), rtol=1e-10)
  np.testing.assert_allclose(odiag.reduce_variance(x, axis=axis, biased=False), np.var(x, axis=axis, ddof=1),
                             rtol=1e-10)


def test_ess_iid_and_block_correlated():
  rng = np.random.default_rng(0)
  x = rng.standard_normal((5000, 3))
  np.testing.assert_allclose(odiag.effective_sample_size(x, filter_threshold=0.), 5000, rtol=0.1)
  np.testing.assert_allclose(odiag.effective_sample_size(x, filter_threshold=None, filter_beyond_positive_pairs=True),
                             5000, rtol=0.25)
  # each value repeated 10 times -> ESS ~ N/10 (diagnostic_test.py:102-140)
  y = np.repeat(rng.standard_normal((500, 2)), 10, axis=0)
  np.testing.assert_allclose(odiag.effective_sample_size(y, filter_beyond_lag=None, filter_threshold=0.),
                             500, rtol=0.25)


def test_cross_chain_ess_counts_modes():
  # diagnostic_test.py:298-331: 4 chains stuck at 4 well separated modes -> ESS ~ 4
  rng = np.random.default_rng(1)
  x = rng.standard_normal((1000, 4)) * 0.01 + np.array([-30., -10., 10., 30.])
  ess = odiag.effective_sample_size(x, cross_chain_dims=1, filter_beyond_positive_pairs=True)
  np.testing.assert_allclose(ess, 4., rtol=0.06)


def test_autocov_matches_direct_sum():
  rng = np.random.default_rng(2)
  x = rng.standard_normal((50, 2))
  ac = odiag.auto_covariance(x)
  xc = x - x.mean(0)
  for k in (0, 1, 7, 49):
    direct = (xc[:50 - k] * xc[k:]).sum(0) / (50 - k)
    np.testing.assert_allclose(ac[k], direct, rtol=1e-9, atol=1e-12)


# ---- leapfrog energy conservation (leapfrog_integrator_test.py:28-83) ------------------
def test_leapfrog_energy_conservation_pin():
  class LogGamma:   # log-density of log-gamma(alpha=5, beta=10) variable: alpha*x - beta*exp(x)
    part_sizes = [3]

    @staticmethod
    def logp_grad(x):
      x = np.asarray(x, np.float32)
      lp = np.sum(np.float32(5.) * x - np.float32(10.) * np.exp(x), axis=1)
      return lp.astype(np.float32), (np.float32(5.) - np.float32(10.) * np.exp(x)).astype(np.float32)

  # x = 0.1 * normal(shape=(50, 10, 2)), one chain dim, event size 20, step 0.09 / event_size
  rng = np.random.default_rng(0)
  x = (0.1 * rng.standard_normal((50, 20))).astype(np.float32)
  m = rng.standard_normal(x.shape).astype(np.float32)
  lp, g = LogGamma.logp_grad(x)
  e0 = -lp + 0.5 * np.sum(m * m, axis=1)
  eps = np.full(x.shape, 0.09 / 20, np.float32)
  m1, x1, lp1, _ = omcmc.leapfrog(LogGamma, m, x, lp, g, eps, 1000)
  e1 = -lp1 + 0.5 * np.sum(m1 * m1, axis=1)
  assert np.all(np.abs(e1 - e0) / np.abs(e0) <= 0.02)


# ---- gradients vs autodiff (mcmc/internal/util_test.py:163-200 in spirit) --------------
def test_analytic_gradients_vs_torch_autograd():
  torch = pytest.importorskip('torch')
  N = torch.distributions.Normal
  es = otargets.EightSchools(dtype=np.float64)
  x = np.random.default_rng(0).standard_normal((5, 10))
  lp, g = es.logp_grad(x)
  xt = torch.tensor(x, requires_grad=True)
  y, s = torch.tensor(es.y), torch.tensor(es.sigma)
  l = (N(0., 10.).log_prob(xt[:, 0]) + N(5., 1.).log_prob(xt[:, 1]) + N(0., 1.).log_prob(xt[:, 2:]).sum(1)
       + N(xt[:, 0:1] + torch.exp(xt[:, 1:2]) * xt[:, 2:], s).log_prob(y).sum(1))
  l.sum().backward()
  np.testing.assert_allclose(lp, l.detach().numpy(), rtol=1e-6)
  np.testing.assert_allclose(g, xt.grad.numpy(), rtol=1e-9, atol=1e-12)

  T = 40
  yv = otargets.synthetic_sv_returns(T=T).astype(np.float64)
  sv = otargets.StochasticVolatility(yv, dtype=np.float64)
  u = np.random.default_rng(1).standard_normal((4, T + 3)) * 0.5
  lp, g = sv.logp_grad(u)
  ut = torch.tensor(u, requires_grad=True)
  sp = torch.nn.functional.softplus
  phi = 2 * torch.sigmoid(ut[:, 0]) - 1; m = ut[:, 1]; sc = sp(ut[:, 2]); z = ut[:, 3:]
  h = [sc * z[:, 0] / torch.sqrt(1 - phi ** 2)]
  for i in range(1, T):
    h.append(phi * h[-1] + sc * z[:, i])
  h = torch.stack(h, 1)
  f64 = lambda v: torch.tensor(v, dtype=torch.float64)
  lik = N(f64(0.), torch.exp(0.5 * (h + m[:, None]))).log_prob(torch.tensor(yv)).sum(1)
  b = (phi + 1) / 2
  l = (lik + torch.distributions.Beta(f64(20.), f64(1.5)).log_prob(b) - np.log(2.)
       + torch.distributions.Cauchy(f64(0.), f64(5.)).log_prob(m)
       + torch.distributions.HalfCauchy(f64(2.)).log_prob(sc) + N(f64(0.), f64(1.)).log_prob(z).sum(1)
       + np.log(2.) - sp(-ut[:, 0]) - sp(ut[:, 0]) - sp(-ut[:, 2]))
  l.sum().backward()
  np.testing.assert_allclose(lp, l.detach().numpy(), rtol=1e-10)
  np.testing.assert_allclose(g, ut.grad.numpy(), rtol=1e-8, atol=1e-10)

  X, yy = otargets.synthetic_logistic_data(50, 4)
  lg = otargets.LogisticRegression(X.astype(np.float64), yy.astype(np.float64), dtype=np.float64)
  th = np.random.default_rng(2).standard_normal((3, 5))
  lp, g = lg.logp_grad(th)
  tt = torch.tensor(th, requires_grad=True)
  zz = tt @ torch.tensor(X.astype(np.float64)).T
  l = N(f64(0.), f64(1.)).log_prob(tt).sum(1) + torch.distributions.Bernoulli(logits=zz).log_prob(
      torch.tensor(yy.astype(np.float64))).sum(1)
  l.sum().backward()
  np.testing.assert_allclose(lp, l.detach().numpy(), rtol=1e-10)
  np.testing.assert_allclose(g, tt.grad.numpy(), rtol=1e-9, atol=1e-12)

  cov, _ = otargets.ill_conditioned_covariance(10)
  P, c = otargets.gaussian_precision_from_cov(cov)
  dg = otargets.DenseGaussian(P.astype(np.float64), c, dtype=np.float64)
  xx = np.random.default_rng(3).standard_normal((4, 10))
  lp, g = dg.logp_grad(xx)
  mvn = torch.distributions.MultivariateNormal(torch.zeros(10, dtype=torch.float64),
                                               covariance_matrix=torch.tensor(cov))
  np.testing.assert_allclose(lp, mvn.log_prob(torch.tensor(xx)).numpy(), rtol=2e-4)


def test_ill_conditioned_gaussian_spec():
  cov, ev = otargets.ill_conditioned_covariance(100)
  assert 1e5 < ev.max() / ev.min() < 2e5          # cond ~ 1.3e5
  np.testing.assert_allclose(np.sort(np.linalg.eigvalsh(cov)), np.sort(ev), rtol=1e-6)


# ---- NUTS: lock-step batched form == per-chain early-exit form --------------------------
@pytest.mark.parametrize('layout', [0, 1])
def test_nuts_lockstep_equals_per_chain(layout):
  es = otargets.EightSchools()
  B = 12
  x0 = np.tile(np.array([0, 0] + [1] * 8, np.float32), (B, 1)) + \
      0.1 * np.random.default_rng(3).standard_normal((B, 10)).astype(np.float32)
  lp0, g0 = es.logp_grad(x0)
  for seed in (1, 2):
    k = orng.key(seed)
    a = omcmc.nuts_one_step(es, x0, lp0, g0, 0.3, k, max_tree_depth=6, layout=layout)
    b = omcmc.nuts_one_step_chain(es, x0, lp0, g0, 0.3, k, max_tree_depth=6, layout=layout)
    for f in a:
      np.testing.assert_array_equal(a[f], b[f], err_msg=f)
  assert len(set(a['leapfrogs_taken'].tolist())) > 1      # ragged trees were exercised


def test_nuts_sharded_chains_equal_unsharded():
  """RNG counters use the global chain index: shards reproduce the unsharded run bit for bit."""
  es = otargets.EightSchools()
  B = 8
  x0 = np.tile(np.array([0, 0] + [1] * 8, np.float32), (B, 1))
  x0 += 0.1 * np.random.default_rng(4).standard_normal((B, 10)).astype(np.float32)
  lp0, g0 = es.logp_grad(x0)
  k = orng.key(5)
  full = omcmc.nuts_one_step(es, x0, lp0, g0, 0.3, k, max_tree_depth=5)
  for off in (0, 4):
    sh = omcmc.nuts_one_step(es, x0[off:off + 4], lp0[off:off + 4], g0[off:off + 4], 0.3, k, max_tree_depth=5,
                             chain_offset=off, B_global=B)
    for f in full:
      np.testing.assert_array_equal(full[f][off:off + 4], sh[f], err_msg=f)
  h_full = omcmc.hmc_one_step(es, x0, lp0, g0, 0.3, 3, k)
  h_sh = omcmc.hmc_one_step(es, x0[4:], lp0[4:], g0[4:], 0.3, 3, k, chain_offset=4, B_global=B)
  np.testing.assert_array_equal(h_full['state'][4:], h_sh['state'])


# ---- statistical pins on the oracle samplers (hmc_test.py:142-200, nuts_test.py:310-356) --
def test_oracle_hmc_samples_standard_normal():
  dg = otargets.DenseGaussian(np.eye(2, dtype=np.float32), -np.log(2 * np.pi))
  x0 = np.zeros((256, 2), np.float32)
  states, trace, _ = omcmc.sample_chain(dg, 'hmc', x0, 200, 50, 0, step_size=0.9, num_leapfrog_steps=3, seed=3)
  s = states.reshape(-1, 2)
  assert np.all(np.abs(s.mean(0)) < 0.05)
  assert np.all(np.abs(s.var(0) - 1) < 0.08)
  acc = np.mean([t['is_accepted'].mean() for t in trace])
  assert 0.6 < acc <= 1.0


def test_oracle_nuts_samples_correlated_normal():
  cov = np.array([[1.0, 0.6], [0.6, 2.0]])
  P, c = otargets.gaussian_precision_from_cov(cov)
  dg = otargets.DenseGaussian(P, c)
  x0 = np.zeros((128, 2), np.float32)
  states, trace, _ = omcmc.sample_chain(dg, 'nuts', x0, 150, 50, 0, step_size=0.5, max_tree_depth=5, seed=4)
  s = states.reshape(-1, 2)
  assert np.all(np.abs(s.mean(0)) < 0.1)
  np.testing.assert_allclose(np.cov(s.T), cov, atol=0.15)
  assert all((t['leapfrogs_taken'] >= 1).all() for t in trace)
