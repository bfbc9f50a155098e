"""CPU: the oracle against the committed golden fixtures (tests/golden/, generated from the reference
by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

from oracle import mcmc as omcmc
from oracle import targets as otargets

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_nuts_tables_vs_golden():
  g = json.load(open(os.path.join(G, 'nuts_tables_depth4.json')))
  for fn in (omcmc.write_read_instructions, omcmc.write_read_closed_form):
    w, r = fn(g['max_tree_depth'])
    np.testing.assert_array_equal(w, g['write_instruction'])
    np.testing.assert_array_equal(r, g['read_instruction'])


def test_dual_averaging_vs_golden():
  g = json.load(open(os.path.join(G, 'dual_averaging_pins.json')))
  for name, err in (('_UPDATE_M05', -0.05), ('_UPDATE_M02', -0.02), ('_UPDATE_M01', -0.01), ('_UPDATE_0', 0.0),
                    ('_UPDATE_01', 0.01)):
    da = omcmc.DualAveraging(1.0, 100, step_count_smoothing=g['_INITIAL_T'],
                             exploration_shrinkage=g['_EXPLORATION_SHRINKAGE'])
    got = da.update(np.log(np.full(2, 0.75 + err, np.float32)))
    np.testing.assert_allclose(got, g[name], rtol=2e-6)


def test_sv_fixture_and_oracle_at_ground_truth_mean():
  """The S&P 500 fixture has the reference's shape (2516 centred returns) and the oracle's
  unconstrained log-density is finite with a finite gradient at the Stan posterior mean."""
  z = np.load(os.path.join(G, 'sv_sp500.npz'))
  y = z['centered_returns']
  assert y.shape == (2516,) and abs(y.mean()) < 1e-9
  sv = otargets.StochasticVolatility(y.astype(np.float64), dtype=np.float64)
  phi = float(z['identity_persistence_of_volatility_mean'])
  m = float(z['identity_mean_log_volatility_mean'])
  s = float(z['identity_white_noise_shock_scale_mean'])
  assert 0.9 < phi < 1 and 0 < s < 1
  h = z['identity_log_volatility_mean'] - m        # centred log-vol at the posterior mean
  zz = np.empty(2516)
  zz[0] = h[0] * np.sqrt(1 - phi * phi) / s
  zz[1:] = (h[1:] - phi * h[:-1]) / s
  u = np.concatenate([[np.log((phi + 1) / (1 - phi)), m, np.log(np.expm1(s))], zz])[None, :]
  lp, g = sv.logp_grad(u)
  assert np.isfinite(lp).all() and np.isfinite(g).all()
  # constrain() inverts the parameterisation used above
  p2, m2, s2, _ = sv.constrain(u)
  np.testing.assert_allclose([p2[0], m2[0], s2[0]], [phi, m, s], rtol=1e-9)
