"""tfp.experimental surface of the hot path: preconditioned (diagonal mass matrix) HMC / NUTS, mass-matrix and
windowed adaptation, streaming reducers (tfp/experimental/mcmc, tfp/experimental/stats)."""
from probability_b200.experimental import mcmc
from probability_b200.experimental import stats
