/* pb2.h -- C ABI of libpb2 (probability_b200): the B200-native many-chain HMC/NUTS
 * engine behind the tfp.mcmc hot path.
 *
 * The reference (TensorFlow Probability) has NO native/FFI boundary for this path:
 * it is Python over TF/JAX ops.  The interfaces replaced here are therefore the
 * reference's Python operator API (paths relative to tensorflow_probability/python/):
 *
 *   pb2_rng_*            internal/samplers.py:79-368 (sanitize_seed, fold_in, split_seed,
 *                        normal, uniform) over jax.random
 *                        (internal/backend/numpy/random_generators.py:151-158,278-302)
 *   pb2_logp_grad        mcmc/internal/util.py:286-308 maybe_call_fn_and_grads
 *   pb2_dense_/pb2_logistic_logp_grad_tc   the same, all chains at once on the tcgen05 tensor cores
 *   pb2_rowshard_*       target_log_prob_fn + gradient of a logistic regression whose rows are sharded over ranks
 *                        (psum in the target: internal/distribute_lib.py:147-162,179-242)
 *   pb2_leapfrog         mcmc/internal/leapfrog_integrator.py:222-316 SimpleLeapfrogIntegrator.__call__
 *   pb2_run (HMC)        mcmc/hmc.py:501-529,661-729 + mcmc/metropolis_hastings.py:160-254
 *   pb2_run (NUTS)       mcmc/nuts.py:321-445 NoUTurnSampler.one_step
 *   pb2_run (n_steps>1)  mcmc/sample.py:81-383 sample_chain (+ internal/loop_util.py:123-253)
 *   pb2_da_*             mcmc/dual_averaging_step_size_adaptation.py:353-532
 *   pb2_comm_*           internal/distribute_lib.py:147-242 (named-axis psum / reduce_logsumexp) -> NCCL
 *   pb2_rowshard_leapfrog  leapfrog_integrator.py:280-355 over a psum'd target (distribute_lib.py:179-242)
 *   pb2_ess / pb2_rhat   mcmc/diagnostic.py:38-336,339-567
 *
 * Conventions: every function returns 0 on success and a negative code on error
 * (message via pb2_last_error); nothing throws or aborts.  All array arguments named
 * d_* are DEVICE pointers (float32 / int32 / uint8, row-major, contiguous) owned by
 * the caller; h_* are HOST pointers.  Work is enqueued on the context's CUDA stream
 * (pb2_ctx_set_stream) and is asynchronous unless the function returns host values.
 * One context per (host thread, GPU); calls on one context are not re-entrant.
 */
#ifndef PB2_H_
#define PB2_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB2_VERSION 100

#define PB2_OK 0
#define PB2_ERR_INVALID (-1)
#define PB2_ERR_CUDA (-2)
#define PB2_ERR_UNSUPPORTED (-3)

/* RNG counter layouts (jax_threefry_partitionable True / False). */
#define PB2_LAYOUT_PARTITIONABLE 0
#define PB2_LAYOUT_ORIGINAL 1
/* Philox-4x32-10 bit generator (the generator of tf.random.stateless_* behind samplers.py:249-250,324-325 on the TF
 * substrate): element j of a flat draw = word j % 4 of the block at counter j / 4 under the uint32[2] key;
 * split(key, n) = bits(2n) reshaped; fold_in and the float transforms are those of the threefry layouts */
#define PB2_LAYOUT_PHILOX 2

/* target kinds */
#define PB2_TARGET_EIGHT_SCHOOLS 0
#define PB2_TARGET_DENSE_GAUSSIAN 1
#define PB2_TARGET_LOGISTIC 2
#define PB2_TARGET_STOCH_VOL 3
#define PB2_TARGET_STOCH_VOL_CONSTRAINED 4 /* the same model in its own coordinates [persistence, mean, scale, z] */
#define PB2_TARGET_USER 5 /* user-defined: CUDA source compiled at run time, see pb2_target_create_user */
/* the CENTRED stochastic-volatility model (inference_gym/targets/stochastic_volatility.py:39-111): one latent
 * log-volatility per time step, state [phi, m, s, x[T]]; 6 = unconstrained coordinates with the Sigmoid(-1,1) / Softplus
 * bijectors folded in, 7 = the model's own coordinates (for a TransformedTransitionKernel) */
#define PB2_TARGET_STOCH_VOL_CENTERED 6
#define PB2_TARGET_STOCH_VOL_CENTERED_CONSTRAINED 7

/* elementwise event-space bijectors (one per state dimension), tfp/bijectors/{identity,exp,softplus,sigmoid}.py */
#define PB2_BIJECTOR_IDENTITY 0
#define PB2_BIJECTOR_EXP 1
#define PB2_BIJECTOR_SOFTPLUS 2
#define PB2_BIJECTOR_SIGMOID 3 /* Sigmoid(low, high) */

/* transition kinds */
#define PB2_KERNEL_HMC 0
#define PB2_KERNEL_NUTS 1

/* step size kinds */
#define PB2_STEP_SCALAR 0
#define PB2_STEP_PER_DIM 1
#define PB2_STEP_PER_CHAIN 2

typedef struct pb2_ctx pb2_ctx;
typedef struct pb2_target pb2_target;

/* ---- lifecycle -------------------------------------------------------------- */
int pb2_version(void);
int pb2_ctx_create(int device, pb2_ctx** out);
int pb2_ctx_destroy(pb2_ctx* ctx);
int pb2_ctx_set_stream(pb2_ctx* ctx, void* cuda_stream);
int pb2_ctx_synchronize(pb2_ctx* ctx);
const char* pb2_last_error(pb2_ctx* ctx); /* ctx may be NULL: last global error */
/* number of pb2 kernels launched on this context since creation (bench "gpu_launches") */
long long pb2_launch_count(pb2_ctx* ctx);
/* tuning / A-B switches.  "dense_variant" (dense-Gaussian target): 0 = default (tcgen05 tile kernels where
 * supported -- 32 < D <= 100, B >= 256: 128-chain HMC tiles, 64-chain asynchronous-lane NUTS tiles -- else
 * warp-per-chain), 1 = force the warp-per-chain FP32-FMA kernels, 2 = half-warp-per-chain experiment,
 * 3 = tile kernels with the lock-step NUTS kernel (the reference's literal batched algorithm; the bit-exact
 * partner of the asynchronous-lane kernel in the tests).
 * "rowshard_collective": how pb2_rowshard_leapfrog sums the gradient over ranks, see there (default 1).
 * "synchronous_run": 1 = pb2_run waits for the stream before it returns; default 0 = all work is enqueued on the context's
 * stream and the call returns (host outputs -- the pass-along seed, the step seeds -- are final on return; device outputs
 * are stream-ordered). */
int pb2_ctx_set_int(pb2_ctx* ctx, const char* name, int value);

/* ---- targets ---------------------------------------------------------------- */
typedef struct {
  int kind;       /* PB2_TARGET_* */
  int dim;        /* state dimension D (eight schools: J+2; logistic: incl. bias; SV: T+3) */
  int n_rows;     /* logistic: N; stochastic volatility: T; eight schools: J */
  const float* h_a; /* eight schools: y[J]; dense: precision P[D,D]; logistic: X~[N,D]; SV: returns y[T] */
  const float* h_b; /* eight schools: sigma[J]; dense: loc[D] or NULL; logistic: labels y[N] (0/1) */
  float scalar;   /* dense: log-normaliser constant added to the log-prob */
} pb2_target_desc;

int pb2_target_create(pb2_ctx* ctx, const pb2_target_desc* desc, pb2_target** out);
/* User-defined target = the reference's arbitrary `target_log_prob_fn` (tfp/mcmc/hmc.py:413-415,
 * mcmc/internal/util.py:246-308 value + gradient; inference_gym model_contract.md `unnormalized_log_prob`).  A
 * persistent kernel cannot call back into the host, so the target is CUDA C++ source defining, for ONE chain,
 *     __device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data);
 * (returns log p(x) up to a constant, writes the gradient into g[0..dim)).  It is compiled with NVRTC for sm_100a
 * together with csrc/pb2_user_target.cuh (`include_dir` = the directory holding the pb2 device headers) into the same
 * leapfrog / HMC / NUTS chain kernels the named targets use; 1 <= dim <= 256; `h_data` (n_data floats, may be NULL) is
 * copied to the device and handed to every call.  With PB2_USER_COOPERATIVE in `flags` the function takes a trailing
 * `int lane` and is called by all 32 lanes of the chain's warp (x, g in shared memory; every lane writes a disjoint
 * part of g and all return the same value).  On a compile error the call fails with PB2_ERR_INVALID and
 * pb2_last_error returns the compiler's log.  pb2_user_target_check only compiles (no device needed). */
#define PB2_USER_COOPERATIVE 1
int pb2_target_create_user(pb2_ctx* ctx, int dim, const char* cuda_source, int flags, const float* h_data,
                           long long n_data, const char* include_dir, pb2_target** out);
int pb2_user_target_check(const char* cuda_source, int dim, int flags, const char* include_dir);
int pb2_target_destroy(pb2_target* tgt);
int pb2_target_dim(const pb2_target* tgt);

/* ---- RNG (host-side key algebra; tiny, pure C) --------------------------------- */
int pb2_rng_split(const uint32_t key[2], int n, int layout, uint32_t* h_out /*[n,2]*/);
int pb2_rng_fold_in(const uint32_t key[2], uint32_t data, uint32_t out[2]);
/* ---- RNG (device bulk draws of n elements, flat row-major counters) ------------ */
int pb2_rng_bits(pb2_ctx* ctx, const uint32_t key[2], long long n, int layout, uint32_t* d_out);
int pb2_rng_uniform(pb2_ctx* ctx, const uint32_t key[2], long long n, float lo, float hi, int layout,
                    float* d_out);
int pb2_rng_normal(pb2_ctx* ctx, const uint32_t key[2], long long n, int layout, float* d_out);
int pb2_rng_randint(pb2_ctx* ctx, const uint32_t key[2], long long n, int lo, int hi, int layout,
                    int32_t* d_out);

/* ---- chain layout shared by the calls below ------------------------------------ */
typedef struct {
  int B;             /* chains held by this process */
  int B_global;      /* chains of the whole job (RNG draw shapes use this) */
  int chain_offset;  /* global index of local chain 0 */
  int rng_layout;    /* PB2_LAYOUT_* */
  int n_parts;       /* state parts (one momentum key per part), <= 8 */
  int part_sizes[8]; /* sizes sum to D */
} pb2_chain_layout;

/* ---- primitives (parity surface) ----------------------------------------------- */
int pb2_logp_grad(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_x /*[B,D]*/,
                  float* d_logp /*[B]*/, float* d_grad /*[B,D]*/);

/* pb2_logp_grad of the target composed with event-space bijectors (see pb2_run_cfg.d_bijector_*): d_x is the
 * unconstrained state, logp includes the forward log-det-Jacobian (transformed_kernel.py:86-140). */
int pb2_logp_grad_transformed(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_x, const int32_t* d_bijector_kind,
                              const float* d_bijector_low, const float* d_bijector_high, float* d_logp, float* d_grad);

int pb2_leapfrog(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_m, const float* d_x,
                 const float* d_logp, const float* d_grad, const float* d_step, int step_kind,
                 int num_steps, float* d_m_out, float* d_x_out, float* d_logp_out, float* d_grad_out);

/* Same contract as pb2_logp_grad for a dense-Gaussian target (D <= 100), evaluated for all chains at once
 * on the tcgen05 tensor cores (128-chain tiles, 3xTF32 split, accumulator in TMEM).  pb2_logp_grad
 * dispatches here for B >= 128. */
int pb2_dense_logp_grad_tc(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_x, float* d_logp,
                           float* d_grad);

/* Same contract as pb2_logp_grad for a logistic-regression target (D <= 32 incl. the bias column), evaluated for
 * all chains at once on the tcgen05 tensor cores: logits = Theta X~^T and grad = (y - sigmoid(logits)) X~ as two
 * 3xTF32 contractions per 128-chain tile with the [128 x N] logits kept in TMEM (gym logistic_regression.py:88-103). */
int pb2_logistic_logp_grad_tc(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_x, float* d_logp,
                              float* d_grad);

/* ---- transitions / sample_chain -------------------------------------------------- */
typedef struct {
  int kind;                 /* PB2_KERNEL_HMC | PB2_KERNEL_NUTS */
  int num_leapfrog_steps;   /* HMC */
  int max_tree_depth;       /* NUTS (<= 12) */
  float max_energy_diff;    /* NUTS */
  int unrolled_leapfrog_steps; /* NUTS */
  int num_results;          /* R */
  int num_burnin_steps;
  int num_steps_between_results;
  int step_kind;            /* PB2_STEP_* */
  int explicit_step_seeds;  /* 0: derive per-transition seeds from h_seed (sample_chain);
                               1: h_step_seeds [n_steps,2] is an INPUT (TransitionKernel.one_step) */
  const float* d_momentum_scale; /* NULL, or [D] device: s = sqrt of the diagonal INVERSE mass matrix (the running
                               variance): momentum ~ N(0, diag(1/s^2)), kinetic energy 1/2 sum s^2 m^2, positions move
                               along the velocity s^2 m, U-turn test on <rho, velocity> -- PreconditionedHamiltonianMonteCarlo /
                               PreconditionedNoUTurnSampler with the momentum distribution DiagonalMassMatrixAdaptation
                               builds (experimental/mcmc/preconditioned_hmc.py, preconditioned_nuts.py:169,
                               diagonal_mass_matrix_adaptation.py:73).  States, gradients and momenta cross the ABI in
                               the ORIGINAL coordinates; inside, the kernels run on u = x / s (pb2_targets.cuh). */
  /* NULL, or [D] device arrays: event-space bijectors of a TransformedTransitionKernel (mcmc/transformed_kernel.py:
   * 86-140,369-419).  The chain state d_x is then the UNCONSTRAINED state, the target is evaluated at
   * forward(state) and the forward log-det-Jacobian and its derivative are added (PB2_BIJECTOR_* per dimension,
   * low / high read for PB2_BIJECTOR_SIGMOID only). */
  const int32_t* d_bijector_kind;
  const float* d_bijector_low;
  const float* d_bijector_high;
} pb2_run_cfg;

/* Nullable per-result outputs; leading dimension R = num_results.
 * Field names follow MetropolisHastingsKernelResults (metropolis_hastings.py:41-56),
 * UncalibratedHamiltonianMonteCarloKernelResults (hmc.py:40-57) and
 * NUTSKernelResults (nuts.py:74-91). */
typedef struct {
  float* d_states;                    /* [R,B,D] */
  float* d_target_log_prob;           /* [R,B] */
  float* d_grads_target_log_prob;     /* [R,B,D] */
  float* d_log_accept_ratio;          /* [R,B] */
  uint8_t* d_is_accepted;             /* [R,B] */
  float* d_step_size;                 /* [R] scalar step size used by the emitting transition */
  float* d_proposed_state;            /* HMC [R,B,D] */
  float* d_proposed_target_log_prob;  /* HMC [R,B] */
  float* d_proposed_grads;            /* HMC [R,B,D] */
  float* d_log_acceptance_correction; /* HMC [R,B] */
  float* d_initial_momentum;          /* HMC [R,B,D] */
  float* d_final_momentum;            /* HMC [R,B,D] */
  int32_t* d_leapfrogs_taken;         /* NUTS [R,B] */
  uint8_t* d_has_divergence;          /* NUTS [R,B] */
  uint8_t* d_reach_max_depth;         /* NUTS [R,B] */
  float* d_energy;                    /* NUTS [R,B] */
} pb2_trace;

/* Dual-averaging state (device resident, 16 floats): see pb2_da_init. */
typedef struct {
  int enabled;                /* 0: fixed step size */
  float* d_state;             /* [16] device, from pb2_da_init */
  int reduce_over_ranks;      /* 1: the chains are sharded over the ranks of the context's communicator (pb2_comm_init)
                                 and the accept statistic is reduced over ALL of them inside pb2_run: one 2-float
                                 all-gather per adapting transition, enqueued on the stream (no host round trip) --
                                 experimental_reduce_chain_axis_names, dual_averaging_step_size_adaptation.py:259-261 */
} pb2_da;

/* Runs num_burnin + 1 + (R-1)*(1+thin) transitions for all B chains, chaining
 * `step_seed, seed = split(seed)` per transition (sample.py:344-349).  h_seed is
 * the already-salted seed on entry and the pass-along seed on return; h_step_seeds
 * (nullable, [n_steps,2]) receives every transition's seed.  d_x/d_logp/d_grad hold
 * the chain state in and out.  d_step_size: scalar | [D] | [B] per cfg.step_kind; with
 * dual averaging enabled (scalar only) it is updated in place after every transition
 * and the reduction over chains spans this process's chains, or -- da->reduce_over_ranks with a
 * communicator attached (pb2_comm_init) -- the chains of all ranks (every rank must make the same call).
 * d_leapfrog_total (nullable, uint64[B]) accumulates gradient evaluations. */
int pb2_run(pb2_ctx* ctx, const pb2_target* tgt, const pb2_chain_layout* layout,
            const pb2_run_cfg* cfg, uint32_t h_seed[2], uint32_t* h_step_seeds, float* d_x,
            float* d_logp, float* d_grad, float* d_step_size, const pb2_da* da,
            const pb2_trace* trace, unsigned long long* d_leapfrog_total);

/* ---- dual averaging ---------------------------------------------------------------- */
/* state layout: [0] error_sum [1] log_averaging_step [2] log_shrinkage_target
 * [3] step (as float) [4] num_adaptation_steps [5] target_accept_prob
 * [6] exploration_shrinkage [7] step_count_smoothing [8] decay_rate [9] step_size */
int pb2_da_init(pb2_ctx* ctx, float step_size, int num_adaptation_steps, float target_accept_prob,
                float exploration_shrinkage, float step_count_smoothing, float decay_rate,
                float log_shrinkage_target /* NaN: log(10*step_size) */, int step, float error_sum,
                float log_averaging_step, float* d_state /*[16]*/);
/* the 8-byte partial of B chains: the sum of their accept probabilities exp(min(0, finite_or(-inf)(log_accept_ratio)))
 * as a little-endian 64-bit fixed-point integer in 2^-36 units (two 32-bit words in the float slots).  Integer sums are
 * associative: partials of any split of the chains over launches or ranks combine to the same bits
 * (reduce_logmeanexp over named axes, math/generic.py:221-274, distribute_lib.py:147-162). */
int pb2_da_partial(pb2_ctx* ctx, const float* d_log_accept_ratio, int B, float* d_partial /*[2]*/);
/* add n_partials partials, update the state with their mean over B_global chains and write the new step size */
int pb2_da_apply(pb2_ctx* ctx, const float* d_partials /*[n,2]*/, int n_partials, long long B_global,
                 float* d_state, float* d_step_size_out /* nullable */);

/* ---- streaming moments (experimental/stats/sample_stats.py RunningVariance; the reducer behind
 * DiagonalMassMatrixAdaptation and experimental/mcmc/with_reductions.py VarianceReducer) --------------------- */
/* d_state [1 + 2 D] = (count, mean[D], sum of squared deviations[D]), zero-initialised by the caller; every row of
 * d_x [rows, D] is one new observation (chains and draws alike).  variance = state[1 + D + d] / count. */
int pb2_running_moments_update(pb2_ctx* ctx, const float* d_x, long long rows, int D, float* d_state);

/* ---- diagnostics ------------------------------------------------------------------- */
/* states [N,B,D] -> ess: cross_chain ? [D] : [B,D].  filter_threshold NaN = None;
 * filter_beyond_lag < 0 = None. */
int pb2_ess(pb2_ctx* ctx, const float* d_states, int N, int B, int D, float filter_threshold,
            int filter_beyond_lag, int filter_beyond_positive_pairs, int cross_chain, float* d_out);
/* states [N,B,D] -> rhat [D] */
int pb2_rhat(pb2_ctx* ctx, const float* d_states, int N, int B, int D, int split_chains, float* d_out);

/* ---- lock-step path for ROW-SHARDED data (BASELINE config 5) ------------------------------
 * The reference expresses this as a target whose log-likelihood is psum'd over a data axis
 * (internal/distribute_lib.py:76-80,179-242; hmc_test.py:1223-1254) inside
 * SimpleLeapfrogIntegrator; here every rank evaluates its rows for ALL chains at once, the
 * caller all-reduces the packed result (NCCL over NVLink) and finishes with the prior. */
/* X~ local rows [N, DP] (rows zero-padded to DP in {32,64,100} floats), labels y[N] (0/1),
 * theta [B, D] -> d_packed [B, D+1] = (X~^T (y - sigmoid(X~ theta)) | sum_n log-lik). */
int pb2_rowshard_logistic_grad(pb2_ctx* ctx, const float* d_X, const float* d_y, int N, int D, int DP,
                               const float* d_theta, int B, float* d_packed);
/* the same on the tcgen05 tensor cores (D <= 100): logits and gradient as two 3xTF32 contractions per 128-chain
 * tile and 32-row chunk, row segments spread over the SMs, partial sums reduced in a fixed order.  The shard's rows
 * are pre-split ONCE into tensor-core operand planes: the caller allocates pb2_rowshard_tc_planes_bytes(N) bytes,
 * fills them with pb2_rowshard_tc_prepare and passes them to every gradient call. */
long long pb2_rowshard_tc_planes_bytes(int N);
int pb2_rowshard_tc_prepare(pb2_ctx* ctx, const float* d_X, int N, int D, int DP, void* d_planes);
int pb2_rowshard_logistic_grad_tc(pb2_ctx* ctx, const void* d_planes, const float* d_y, int N, int D,
                                  const float* d_theta, int B, float* d_packed);
/* grad = -theta + packed[:, :D]; logp = log N(theta; 0, I) + packed[:, D]
 * (inference_gym logistic_regression.py:88-103, bayesian_model.py:100-102). */
int pb2_rowshard_logistic_finish(pb2_ctx* ctx, const float* d_packed, const float* d_theta, int B, int D,
                                 float* d_grad, float* d_logp);
/* leapfrog pieces on [B,D] arrays (leapfrog_integrator.py:280-309,330-355):
 * mode 0: v = m + (eps/2) g, x += eps v;  mode 1: v += eps g, x += eps v;
 * mode 2: v += eps g, m_out = v - (eps/2) g. */
int pb2_lockstep_leapfrog(pb2_ctx* ctx, int mode, int B, int D, const float* d_step, int step_kind,
                          float* d_v, float* d_x, const float* d_g, const float* d_m_in, float* d_m_out);

/* The Metropolis-Hastings step of a lock-step HMC transition as ONE kernel: log_acceptance_correction
 * = 1/2 (sum m0^2 - sum m1^2) (hmc.py:862-875), log_accept_ratio = safe_sum(lp1, -lp0, correction)
 * (metropolis_hastings.py:204-215, util.py:205-235), u ~ uniform at counter chain_offset + b of a [B_global] draw under
 * accept_key, accept iff log u < ratio (:221-227), and mcmc_util.choose (util.py:103-164) of the state and of every
 * per-chain field of the results (d_prev_*: the previous accepted results' momenta and correction).  Outputs must not
 * alias inputs. */
int pb2_hmc_mh_finish(pb2_ctx* ctx, int B, int D, int B_global, int chain_offset, int rng_layout,
                      const uint32_t accept_key[2], const float* d_m0, const float* d_m1, const float* d_x0,
                      const float* d_lp0, const float* d_g0, const float* d_x1, const float* d_lp1, const float* d_g1,
                      const float* d_prev_m0, const float* d_prev_m1, const float* d_prev_corr, float* d_x_out,
                      float* d_lp_out, float* d_g_out, float* d_m0_out, float* d_m1_out, float* d_corr_acc_out,
                      float* d_corr, float* d_log_accept_ratio, unsigned char* d_is_accepted);

/* SimpleLeapfrogIntegrator (leapfrog_integrator.py:280-355) for the logistic-regression target (D <= 32) with all chains
 * in lock-step on the tensor cores: every leapfrog's log-prob + gradient is one launch of the tcgen05 kernel behind
 * pb2_logistic_logp_grad_tc, with one fused kick + drift kernel between them; all num_steps leapfrogs are enqueued by
 * this call.  Same argument meaning as pb2_leapfrog; outputs must not alias inputs.  (The transition path of
 * HamiltonianMonteCarlo on this target for large batches: a fixed L makes lock-step free of waste.) */
int pb2_logistic_tc_leapfrog(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_m, const float* d_x,
                             const float* d_logp, const float* d_grad, const float* d_step, int step_kind, int num_steps,
                             float* d_m_out, float* d_x_out, float* d_logp_out, float* d_grad_out);

/* ---- multi-GPU group ------------------------------------------------------------------------
 * One NCCL communicator per context (one process per GPU).  Replaces the named-axis collectives of
 * internal/distribute_lib.py:147-242 (psum / reduce_logsumexp / pbroadcast) on this path:
 *   - chain-sharded runs: dual-averaging statistics (pb2_run with da->reduce_over_ranks);
 *   - row-sharded data : the per-leapfrog gradient psum inside pb2_rowshard_leapfrog.
 * Rank 0 creates the id (pb2_comm_unique_id) and hands it to the other ranks by any transport (torch.distributed
 * broadcast, MPI, a file); every rank then calls pb2_comm_init.  NCCL is bound at run time (dlopen). */
#define PB2_COMM_ID_BYTES 128
int pb2_comm_unique_id(void* out_id /* PB2_COMM_ID_BYTES host bytes */);
int pb2_comm_init(pb2_ctx* ctx, int nranks, int rank, const void* nccl_unique_id);
int pb2_comm_destroy(pb2_ctx* ctx);   /* collective (all ranks call it): closes the peer-memory mappings behind a barrier */
int pb2_comm_size(pb2_ctx* ctx);   /* 1 without a communicator */
int pb2_comm_rank(pb2_ctx* ctx);
/* in-place sum over the ranks, enqueued on the context's stream (parity surface of the collective) */
int pb2_comm_allreduce_sum(pb2_ctx* ctx, float* d_buf, long long n);

/* SimpleLeapfrogIntegrator (leapfrog_integrator.py:280-355) for the row-sharded logistic regression, all chains in
 * lock-step, ALL num_steps leapfrogs enqueued by this one call: per leapfrog the local gradient pass
 * (d_planes != NULL: tcgen05, else the FP32 kernel on d_X), the sum of the packed [B, D+1] buffer over the ranks of the
 * context's communicator (reduce_over_ranks != 0) and one fused kernel for prior + kick + drift.  The cross-rank sum is
 * by default done INSIDE that fused kernel over peer memory (pb2_ctx_set_int "rowshard_collective" 1: every rank's
 * packed buffer is mapped into every process with CUDA IPC on first use -- a collective call --, the gradient pass
 * publishes into it, per-rank sequence flags travel over NVLink, the ranks' buffers are added in rank order so replicas
 * stay bit-identical; single node, <= 8 ranks); "rowshard_collective" 0 enqueues an NCCL all-reduce between the two
 * kernels instead.  Same argument meaning as pb2_leapfrog; outputs must not alias inputs. */
int pb2_rowshard_leapfrog(pb2_ctx* ctx, const void* d_planes, const float* d_X, const float* d_y, int N, int D, int DP,
                          int B, const float* d_m, const float* d_x, const float* d_logp, const float* d_grad,
                          const float* d_step, int step_kind, int num_steps, int reduce_over_ranks, float* d_m_out,
                          float* d_x_out, float* d_logp_out, float* d_grad_out);

#ifdef __cplusplus
}
#endif
#endif /* PB2_H_ */
