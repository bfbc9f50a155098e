"""tfp.mcmc surface of the hot path (tfp/mcmc/__init__.py:17-70), backed by libpb2."""
from probability_b200.mcmc.diagnostic import effective_sample_size
from probability_b200.mcmc.diagnostic import potential_scale_reduction
from probability_b200.mcmc.dual_averaging_step_size_adaptation import DualAveragingStepSizeAdaptation
from probability_b200.mcmc.dual_averaging_step_size_adaptation import DualAveragingStepSizeAdaptationResults
from probability_b200.mcmc.hmc import HamiltonianMonteCarlo
from probability_b200.mcmc.hmc import MetropolisHastings
from probability_b200.mcmc.hmc import MetropolisHastingsKernelResults
from probability_b200.mcmc.hmc import UncalibratedHamiltonianMonteCarlo
from probability_b200.mcmc.hmc import UncalibratedHamiltonianMonteCarloKernelResults
from probability_b200.mcmc.kernel import TransitionKernel
from probability_b200.mcmc.nuts import NoUTurnSampler
from probability_b200.mcmc.nuts import NUTSKernelResults
from probability_b200.mcmc.sample import CheckpointableStatesAndTrace
from probability_b200.mcmc.sample import sample_chain
from probability_b200.mcmc.sample import StatesAndTrace
from probability_b200.mcmc.simple_step_size_adaptation import SimpleStepSizeAdaptation
from probability_b200.mcmc._engine import ChainShard
