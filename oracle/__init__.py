"""CPU oracle: a NumPy restatement of the tfp.mcmc HMC/NUTS hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE. Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.  The
product (`probability_b200`) never imports it and has no CPU fallback.

What it restates (reference = /root/reference, TFP 0.26.0-dev; tfp/ =
tensorflow_probability/python/):

* rng.py         tfp/internal/samplers.py:79-368 over the JAX substrate
                 (tfp/internal/backend/numpy/random_generators.py:151-158,278-302).
                 The bit generator itself (threefry2x32, jax.random.*) lives in
                 the third-party, UN-PINNED dependency `jax` (setup.py:110-111),
                 absent from /root/reference and from this image; its published
                 algorithm is restated here.
* targets.py     tfp/mcmc/eight_schools_hmc.py:41-76 and the inference_gym
                 targets (ill_conditioned_gaussian.py:66-81, logistic_regression.py
                 :36-103, vectorized_stochastic_volatility.py:47-99,233-356) with
                 analytic gradients in place of autodiff.
* mcmc.py        leapfrog_integrator.py:222-396, hmc.py:661-875,
                 metropolis_hastings.py:160-288, nuts.py:321-1108,
                 dual_averaging_step_size_adaptation.py:353-609, sample.py:311-383.
* diagnostic.py  mcmc/diagnostic.py:203-336,476-589, stats/sample_stats.py:115-215.

PINNING STATUS
  * NUTS instruction tables, dual-averaging constants, R-hat, _reduce_variance:
    pinned by the reference's own known-answer tests (tests/test_oracle_pins.py).
  * RNG bit stream: pinned against the Random123 Threefry-2x32-20 KATs and on
    the outputs of a LIVE JAX run that the reference itself keeps -- the executed
    cells of examples/jupyter_notebooks/TensorFlow_Probability_on_JAX.ipynb
    (PRNGKey(0), random.split, random.normal on the key and on both children,
    tfd.Normal(0, 1).sample(seed=key); plus multi-element and 2-d draws from
    discussion/examples/TFP_and_Jax.ipynb and Distributed_Inference_with_JAX.ipynb;
    tests/golden/jax_notebook_rng.json, made by tests/golden/make_golden.py).  That covers the block function, the
    original key-split layout and the bits -> uniform -> normal transform; the
    partitionable layout has no reference-held vector (JAX-documented values
    only), and neither jax nor tensorflow is installable here.
  * HMC/NUTS transitions: the reference cannot be executed in this image
    (needs tensorflow or jax).  One reference-held RUN exists and is reproduced
    (tests/golden/tf_notebook_hmc.json: the executed sample_chain cell of
    TFP_Release_Notebook_0_11_0.ipynb, whose dynamics do not depend on the
    generator: leapfrog, accept step and burn-in indexing); beyond it the
    transitions are pinned through the reference's statistical / invariant tests
    restated in tests/.  "parity unpinned" for bit-level trajectories of a
    generator-dependent run.
"""
