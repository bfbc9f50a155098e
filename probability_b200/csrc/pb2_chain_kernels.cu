// Persistent chain kernels: one thread group per chain (warp, or whole CTA for the
// 2519-d stochastic-volatility target), chains pulled from a dynamic queue, each group
// running transitions [t0,t1) of its chain back to back with the state in registers.
// sm_100a only.
#include <algorithm>
#include "pb2_internal.h"

namespace pb2 {

// CTAs per SM of the CTA-per-chain (block group) kernels: a chain's leapfrog is a chain of block-wide scans and
// reductions separated by barriers (latency-bound), so a second resident CTA -- another chain -- fills the gaps
#ifndef PB2_BLOCK_GROUP_CTAS_PER_SM
#define PB2_BLOCK_GROUP_CTAS_PER_SM 2
#endif
template <class Grp>
constexpr int min_ctas_per_sm() { return Grp::kIsBlock ? PB2_BLOCK_GROUP_CTAS_PER_SM : 1; }

template <class Grp, int E, class Tgt, int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, min_ctas_per_sm<Grp>())
chain_kernel(const ChainParams p, const typename Tgt::Params tp, const PrimIO io, const SmemPlan plan) {
  chain_body<Grp, E, Tgt, MODE>(p, tp, io, plan);
}

// Host-side stand-in of the run-time-compiled UserT<WarpG, E> (pb2_user_target.cuh): same shared-memory plan, the kernel
// itself comes from the user's build.
template <int E>
struct JitUserT {
  using Params = UserParams;
  static constexpr bool kCkptInSmem = true;
  static size_t cta_smem_floats(const Params&) { return 0; }
  static size_t group_smem_floats(const Params&) { return 2 * 32 * E; }
};
// which run-time build a target type stands for: -1 = none (offline-compiled), 0 plain, 1 ScaledT, 2 TransformedT
template <class Tgt>
struct JitVariant { static constexpr int value = -1; };
template <int E>
struct JitVariant<JitUserT<E>> { static constexpr int value = 0; };
template <class Grp, int E>
struct JitVariant<ScaledT<Grp, E, JitUserT<E>>> { static constexpr int value = 1; };
template <class Grp, int E>
struct JitVariant<TransformedT<Grp, E, JitUserT<E>>> { static constexpr int value = 2; };

template <class Grp, int E, class Tgt, int MODE, int MAXT = 512>
static int launch_t(pb2_ctx* ctx, const typename Tgt::Params& tp, ChainParams& p, const PrimIO& io,
                    const pb2_target* jit = nullptr) {
  SmemPlan plan{};
  auto r4 = [](size_t v) { return (int)((v + 3) & ~size_t(3)); };
  plan.cta_floats = r4(Tgt::cta_smem_floats(tp));
  plan.red_floats = Grp::kIsBlock ? r4(2 * 8 * (Grp::G / 32)) : 0;
  plan.tgt_off = 64;
  int gf = 64 + r4(Tgt::group_smem_floats(tp));
  plan.ck_floats = 0;
  plan.ck_off = -1;
  if (MODE == kModeNUTS) {
    plan.ck_floats = p.max_depth * E * Grp::G;
    if (Tgt::kCkptInSmem) {
      plan.ck_off = gf;
      gf += 2 * plan.ck_floats;
    }
  }
  plan.cold_off = -1;
  if (MODE == kModeNUTS && Grp::kColdInSmem) {
    plan.cold_off = gf;
    gf += Chain<Grp, E, Tgt>::kColdVectors * E * Grp::G;
  }
  plan.group_floats = gf;
  const size_t cta_bytes = 4ull * (plan.cta_floats + plan.red_floats);
  const size_t grp_bytes = 4ull * gf;
  int warps, threads, grid;
  if (Grp::kIsBlock) {
    threads = Grp::G;
    warps = 1;  // groups per CTA
    grid = std::min(p.B, ctx->num_sms * min_ctas_per_sm<Grp>());
  } else {
    if (cta_bytes + grp_bytes > (size_t)ctx->max_smem_optin)
      return set_error(ctx, PB2_ERR_UNSUPPORTED, "target data does not fit in shared memory");
    int wmax = (int)std::min<size_t>(MAXT / Grp::G, (ctx->max_smem_optin - cta_bytes) / grp_bytes);
    // small batches: spread chains over SMs (latency-bound); big batches: fill each SM
    int want = (p.B + ctx->num_sms - 1) / ctx->num_sms;
    warps = std::max(1, std::min(wmax, want));          // groups per CTA
    if (Grp::G < 32) warps = ((warps * Grp::G + 31) / 32) * (32 / Grp::G);   // whole warps
    threads = warps * Grp::G;
    grid = std::min((p.B + warps - 1) / warps, ctx->num_sms);
  }
  const size_t smem_bytes = cta_bytes + (size_t)warps * grp_bytes;
  if (MODE == kModeNUTS && plan.ck_off < 0) {
    size_t need = (size_t)grid * 2 * plan.ck_floats * sizeof(float);
    if (need > ctx->ckpt_bytes) {
      if (ctx->d_ckpt) cudaFree(ctx->d_ckpt);
      ctx->d_ckpt = nullptr;
      if (int rc = check_cuda(ctx, cudaMalloc(&ctx->d_ckpt, need), "cudaMalloc(ckpt)")) return rc;
      ctx->ckpt_bytes = need;
    }
    p.ckpt_global = ctx->d_ckpt;
  }
  if (int rc = check_cuda(ctx, cudaMemsetAsync(ctx->d_queue, 0, sizeof(int), ctx->stream), "memset(queue)"))
    return rc;
  p.queue = ctx->d_queue;
  if constexpr (JitVariant<Tgt>::value >= 0) {
    if (!jit) return set_error(ctx, PB2_ERR_INVALID, "user target: missing target handle");
    if (int rc = user_target_ensure(ctx, const_cast<pb2_target*>(jit), JitVariant<Tgt>::value)) return rc;
    const void* jit_kernel = jit->user_kernels[JitVariant<Tgt>::value][MODE];
    if (int rc = check_cuda(ctx, cudaFuncSetAttribute(jit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)smem_bytes), "cudaFuncSetAttribute(user kernel)"))
      return rc;
    void* args[4] = {(void*)&p, (void*)&tp, (void*)&io, (void*)&plan};
    if (int rc = check_cuda(ctx, cudaLaunchKernel(jit_kernel, dim3(grid), dim3(threads), args, smem_bytes, ctx->stream),
                            "user chain kernel launch"))
      return rc;
  } else {
    auto kfn = chain_kernel<Grp, E, Tgt, MODE, MAXT>;
    if (int rc = check_cuda(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)smem_bytes), "cudaFuncSetAttribute"))
      return rc;
    kfn<<<grid, threads, smem_bytes, ctx->stream>>>(p, tp, io, plan);
  }
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "chain_kernel launch");
}

template <class Grp, int E, class Tgt, int MAXT = 512>
static int launch_mode(pb2_ctx* ctx, const typename Tgt::Params& tp, int mode, ChainParams& p, const PrimIO& io,
                       const pb2_target* jit = nullptr) {
  switch (mode) {
    case kModeLogpGrad: return launch_t<Grp, E, Tgt, kModeLogpGrad, MAXT>(ctx, tp, p, io, jit);
    case kModeLeapfrog: return launch_t<Grp, E, Tgt, kModeLeapfrog, MAXT>(ctx, tp, p, io, jit);
    case kModeHMC: return launch_t<Grp, E, Tgt, kModeHMC, MAXT>(ctx, tp, p, io, jit);
    case kModeNUTS: return launch_t<Grp, E, Tgt, kModeNUTS, MAXT>(ctx, tp, p, io, jit);
  }
  return set_error(ctx, PB2_ERR_INVALID, "bad mode");
}

// the same launch with the target wrapped for diagonal preconditioning when the run carries a scale
template <class Grp, int E, class Tgt, int MAXT = 512>
static int launch_maybe_scaled(pb2_ctx* ctx, const typename Tgt::Params& tp, int mode, ChainParams& p, const PrimIO& io,
                              const pb2_target* jit = nullptr) {
  if (p.bij_kind) {
    using T = TransformedT<Grp, E, Tgt>;
    typename T::Params bp{tp, BijectorSpec{p.bij_kind, p.bij_lo, p.bij_hi}, p.scale, p.D};
    return launch_mode<Grp, E, T, MAXT>(ctx, bp, mode, p, io, jit);
  }
  if (!p.scale) return launch_mode<Grp, E, Tgt, MAXT>(ctx, tp, mode, p, io, jit);
  using S = ScaledT<Grp, E, Tgt>;
  typename S::Params sp{tp, p.scale, p.D};
  return launch_mode<Grp, E, S, MAXT>(ctx, sp, mode, p, io, jit);
}

int launch_chain(pb2_ctx* ctx, const pb2_target* tgt, int mode, ChainParams& p, const PrimIO& io) {
  const int D = tgt->dim;
  if (tile_path_supported(ctx, tgt, mode, p)) return launch_tile_chain(ctx, tgt, mode, p);
  switch (tgt->kind) {
    case PB2_TARGET_USER: {   // run-time-compiled kernels, same launcher (and the same wrappers) as the named targets
      UserParams tp{tgt->d_a, tgt->n_rows, D};
      switch (user_elements_per_lane(D)) {
        case 1: return launch_maybe_scaled<WarpG, 1, JitUserT<1>>(ctx, tp, mode, p, io, tgt);
        case 2: return launch_maybe_scaled<WarpG, 2, JitUserT<2>>(ctx, tp, mode, p, io, tgt);
        case 4: return launch_maybe_scaled<WarpG, 4, JitUserT<4>>(ctx, tp, mode, p, io, tgt);
        case 8: return launch_maybe_scaled<WarpG, 8, JitUserT<8>>(ctx, tp, mode, p, io, tgt);
      }
      return set_error(ctx, PB2_ERR_UNSUPPORTED, "user target: D > 256 not supported");
    }
    case PB2_TARGET_EIGHT_SCHOOLS: {
      EightSchoolsParams tp{tgt->d_a, tgt->d_b, tgt->n_rows};
      return launch_maybe_scaled<WarpG, 1, EightSchoolsT<WarpG, 1>>(ctx, tp, mode, p, io);
    }
    case PB2_TARGET_DENSE_GAUSSIAN: {
      DenseGaussianParams tp{tgt->d_a, tgt->d_b, tgt->scalar, D};
      if (p.bij_kind) {   // bijectors: the generic wrapper (which also applies the scale) around the plain target
        if (D <= 32) return launch_maybe_scaled<WarpG, 1, DenseGaussianT<WarpG, 1>>(ctx, tp, mode, p, io);
        if (D <= 128) return launch_maybe_scaled<WarpG, 4, DenseGaussianT<WarpG, 4>>(ctx, tp, mode, p, io);
      }
      tp.scale = p.scale;
      if (D <= 32) return launch_mode<WarpG, 1, DenseGaussianT<WarpG, 1>>(ctx, tp, mode, p, io);
      if (D <= 128) {
        // experiment kept for A/B profiling (PB2_DENSE_VARIANT=2): two chains per warp, 16 lanes x 8 elements,
        // one LDS of a P row feeding both -- measured 0.6x of warp-per-chain on B200 (issue-bound, 8 warps/SM)
        if (ctx->dense_variant == 2)
          return launch_mode<HalfWarpG, 8, DenseGaussianT<HalfWarpG, 8>, 256>(ctx, tp, mode, p, io);
        return launch_mode<WarpG, 4, DenseGaussianT<WarpG, 4>>(ctx, tp, mode, p, io);
      }
      return set_error(ctx, PB2_ERR_UNSUPPORTED, "dense Gaussian: D > 128 not supported");
    }
    case PB2_TARGET_LOGISTIC: {
      LogisticParams tp{tgt->d_a, tgt->d_b, tgt->n_rows, D};
      if (D <= 8) return launch_maybe_scaled<WarpG, 1, LogisticT<WarpG, 1, 8>>(ctx, tp, mode, p, io);
      if (D <= 25) return launch_maybe_scaled<WarpG, 1, LogisticT<WarpG, 1, 25>>(ctx, tp, mode, p, io);
      if (D <= 32) return launch_maybe_scaled<WarpG, 1, LogisticT<WarpG, 1, 32>>(ctx, tp, mode, p, io);
      return set_error(ctx, PB2_ERR_UNSUPPORTED, "shared-memory logistic: D > 32 (use the row-sharded path)");
    }
    case PB2_TARGET_STOCH_VOL_CONSTRAINED: {
      StochVolParams tp{tgt->d_a, tgt->n_rows};
      if (D <= 5 * 512) return launch_maybe_scaled<BlockG<16>, 5, StochVolT<BlockG<16>, 5, false>>(ctx, tp, mode, p, io);
      return set_error(ctx, PB2_ERR_UNSUPPORTED, "stochastic volatility: T > 2557 not supported");
    }
    case PB2_TARGET_STOCH_VOL: {
      StochVolParams tp{tgt->d_a, tgt->n_rows};
      if (D <= 5 * 512) return launch_maybe_scaled<BlockG<16>, 5, StochVolT<BlockG<16>, 5>>(ctx, tp, mode, p, io);
      return set_error(ctx, PB2_ERR_UNSUPPORTED, "stochastic volatility: T > 2557 not supported");
    }
    case PB2_TARGET_STOCH_VOL_CENTERED: {
      StochVolParams tp{tgt->d_a, tgt->n_rows};
      if (D <= 5 * 512) return launch_maybe_scaled<BlockG<16>, 5, StochVolCenteredT<BlockG<16>, 5>>(ctx, tp, mode, p, io);
      return set_error(ctx, PB2_ERR_UNSUPPORTED, "stochastic volatility: T > 2557 not supported");
    }
    case PB2_TARGET_STOCH_VOL_CENTERED_CONSTRAINED: {
      StochVolParams tp{tgt->d_a, tgt->n_rows};
      if (D <= 5 * 512)
        return launch_maybe_scaled<BlockG<16>, 5, StochVolCenteredT<BlockG<16>, 5, false>>(ctx, tp, mode, p, io);
      return set_error(ctx, PB2_ERR_UNSUPPORTED, "stochastic volatility: T > 2557 not supported");
    }
  }
  return set_error(ctx, PB2_ERR_INVALID, "unknown target kind");
}

}  // namespace pb2
