"""probability_b200: B200-native many-chain HMC/NUTS engine behind the tfp.mcmc surface.

    import probability_b200 as tfp
    target = tfp.targets.EightSchools()
    kernel = tfp.mcmc.HamiltonianMonteCarlo(target, step_size=0.4, num_leapfrog_steps=3)
    states, kr = tfp.mcmc.sample_chain(1000, current_state=[mu, tau, z], kernel=kernel, seed=17)

Everything that computes runs in hand-written sm_100a CUDA kernels behind the C ABI in
include/pb2.h (libpb2.so, loaded with ctypes).  There is no CPU fallback.
"""
from probability_b200 import bijectors
from probability_b200 import distribute
from probability_b200 import experimental
from probability_b200 import mcmc
from probability_b200 import random
from probability_b200 import targets

__version__ = '0.1.0'
