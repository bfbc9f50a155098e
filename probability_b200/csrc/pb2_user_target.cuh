// User-defined targets (the reference accepts any `target_log_prob_fn` callable, tfp/mcmc/hmc.py:413-415, and
// inference_gym's model contract asks for `unnormalized_log_prob`; spinoffs/inference_gym/model_contract.md).  A
// persistent CUDA kernel cannot call back into Python, so a user target is CUDA C++ SOURCE: one function, plain serial
// code for ONE chain,
//
//     __device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data);
//
// that returns log p(x) (up to a constant) and writes d log p / dx into g[0 .. D) -- the analytic gradient takes the
// place of autodiff (tfp/mcmc/internal/util.py:246-308).  pb2_user.cu compiles this header + the user's source with
// NVRTC for sm_100a at target creation and instantiates the SAME chain kernels the named targets use
// (pb2_chain_kernel.cuh: leapfrog, HMC + Metropolis-Hastings, the iterative NUTS tree, dual averaging around them),
// warp-per-chain, E = ceil(D / 32) state elements per lane.
//
// Execution: the warp gathers the chain's vector into shared memory, lane 0 runs the user's function (SIMT: the other
// lanes idle for exactly as long; no races on g), the lanes pick up their elements of g.  A target created with the
// PB2_USER_COOPERATIVE flag instead provides
//     __device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data, int lane);
// called by all 32 lanes (x, g in shared memory): each lane must write a disjoint part of g and all lanes must return
// the same value (pb2::warp_sum is available).
#pragma once
#include "pb2_chain_kernel.cuh"
#include "pb2_targets.cuh"

namespace pb2 {

struct UserParams {
  const float* data;   // device, n_data floats (may be null)
  int n_data;
  int D;
};

}  // namespace pb2

#ifdef __CUDACC_RTC__
#ifdef PB2_USER_COOPERATIVE
__device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data, int lane);
#else
__device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data);
#endif

namespace pb2 {

template <class Grp, int E>
struct UserT {
  static_assert(!Grp::kIsBlock && Grp::G == 32, "user targets run warp-per-chain");
  using Params = UserParams;
  static constexpr bool kCkptInSmem = true;
  const float* data;
  int n_data, D;
  float* xs;   // [32 E] the chain's vector, gathered
  float* gs;   // [32 E] its gradient
  // shared-memory plan (the host side uses the stand-in JitUserT of pb2_chain_kernels.cu): no CTA-wide data,
  // 2 x 32 E floats per group
  PB2_HOSTFN static size_t cta_smem_floats(const Params&) { return 0; }
  PB2_HOSTFN static size_t group_smem_floats(const Params&) { return 2 * Grp::G * E; }
  __device__ void init_cta(const Params&, float*) {}
  __device__ void init_group(const Params& p, Grp&, float*, float* gsm) {
    data = p.data; n_data = p.n_data; D = p.D;
    xs = gsm;
    gs = gsm + Grp::G * E;
  }
  __device__ float logp_grad(Grp& grp, const float (&x)[E], float (&g)[E]) {
#pragma unroll
    for (int j = 0; j < E; ++j) xs[grp.lane * E + j] = x[j];
    grp.sync();
    float lp = 0.f;
#ifdef PB2_USER_COOPERATIVE
    lp = target_log_prob_and_grad(xs, gs, data, n_data, grp.lane);
#else
    if (grp.lane == 0) lp = target_log_prob_and_grad(xs, gs, data, n_data);
    lp = grp.bcast(lp, 0);
#endif
    grp.sync();
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int d = grp.lane * E + j;
      g[j] = d < D ? gs[d] : 0.f;
    }
    return lp;
  }
};

}  // namespace pb2

#ifndef PB2_USER_E
#error "PB2_USER_E (state elements per lane) must be defined by the run-time build"
#endif
// one build per VARIANT of the target: 0 = as written, 1 = diagonally preconditioned (ScaledT: windowed adaptation,
// PreconditionedHMC / NUTS), 2 = behind event-space bijectors (TransformedT: TransformedTransitionKernel)
#ifndef PB2_USER_VARIANT
#define PB2_USER_VARIANT 0
#endif
namespace pb2 {
using UserBase = UserT<WarpG, PB2_USER_E>;
#if PB2_USER_VARIANT == 0
using UserTgt = UserBase;
#elif PB2_USER_VARIANT == 1
using UserTgt = ScaledT<WarpG, PB2_USER_E, UserBase>;
#else
using UserTgt = TransformedT<WarpG, PB2_USER_E, UserBase>;
#endif
}  // namespace pb2
#define PB2_USER_KERNEL(name, mode)                                                                             \
  extern "C" __global__ void __launch_bounds__(512, 1)                                                          \
  name(const pb2::ChainParams p, const pb2::UserTgt::Params tp, const pb2::PrimIO io, const pb2::SmemPlan plan) { \
    pb2::chain_body<pb2::WarpG, PB2_USER_E, pb2::UserTgt, mode>(p, tp, io, plan);                                \
  }
PB2_USER_KERNEL(pb2_user_logp_grad, pb2::kModeLogpGrad)
PB2_USER_KERNEL(pb2_user_leapfrog, pb2::kModeLeapfrog)
PB2_USER_KERNEL(pb2_user_hmc, pb2::kModeHMC)
PB2_USER_KERNEL(pb2_user_nuts, pb2::kModeNUTS)
#endif  // __CUDACC_RTC__
