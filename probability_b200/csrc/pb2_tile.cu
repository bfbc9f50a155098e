// 128-chain TILE transition kernels for the dense-Gaussian target (tcgen05 path).
//   tile_hmc_kernel : HamiltonianMonteCarlo = MetropolisHastings(UncalibratedHMC)
//                     (tfp/mcmc/hmc.py:661-729,780-875; metropolis_hastings.py:181-254)
// All chains of a tile advance in lock-step; each leapfrog's gradient is one 3xTF32 tensor-core
// contraction (pb2_tile.cuh).  Same seeds, same counters, same decisions as the warp-per-chain
// kernels and the oracle (the uint32 streams are identical; floats agree to rounding).
#include <cstdio>
#include <cstdlib>
#include "pb2_tile_nuts.cuh"

namespace pb2 {
using namespace tile;

struct TileIO {
  float x[kK];
  float g[kK];
};

__device__ __forceinline__ void tile_load(const float* base, int c, int D, int slice, bool live, float (&v)[kK]) {
  const float* row = base + (size_t)c * D + kK * slice;
#pragma unroll
  for (int j = 0; j < kK; ++j) v[j] = (live && kK * slice + j < D) ? row[j] : 0.f;
}

__device__ __forceinline__ void tile_store(float* base, size_t r, int B, int c, int D, int slice, bool live,
                                           const float (&v)[kK]) {
  if (!live) return;
  float* row = base + (r * (size_t)B + c) * D + kK * slice;
#pragma unroll
  for (int j = 0; j < kK; ++j)
    if (kK * slice + j < D) row[j] = v[j];
}

__device__ __forceinline__ int tile_result_index(const ChainParams& p, int t) {
  int u = t - p.burnin;
  if (u < 0) return -1;
  int q = u / (p.thin + 1);
  if (q * (p.thin + 1) != u || q >= p.n_results) return -1;
  return q;
}

__device__ __forceinline__ float tile_eps(const ChainParams& p, int c, bool live) {
  return p.step_kind == 0 ? p.step[0] : (live ? p.step[c] : 0.f);
}

// momentum ~ N(0, I): one key per state part, counter = row-major index in [B_global, size_part]
__device__ __forceinline__ float tile_momentum(const ChainParams& p, const uint32_t* keys, uint64_t cg, int d) {
  int part = 0;
#pragma unroll 1
  for (int q = 1; q < p.n_parts; ++q) part += (d >= p.part_off[q]) ? 1 : 0;
  const int off = p.part_off[part];
  const uint64_t sz = (uint64_t)(p.part_off[part + 1] - off);
  Key k{keys[2 * part], keys[2 * part + 1]};
  return normal_from_bits(bits_at(k, cg * sz + (uint64_t)(d - off), (uint64_t)p.B_global * sz, p.layout));
}

__global__ void __launch_bounds__(kThreads, 1)
tile_hmc_kernel(const ChainParams p, const DenseGaussianParams tp) {
  extern __shared__ __align__(128) unsigned char planes[];
  __shared__ Shared sh;
  Ctx cx;
  cx.init(&sh, planes, tp.P, tp.loc, tp.D);
  const int D = tp.D;
  const int ntiles = (p.B + kM - 1) / kM;
  for (int tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    const int c = tile_i * kM + cx.cl;
    const bool live = c < p.B;
    const uint64_t cg = (uint64_t)p.chain_offset + (uint64_t)c;
    float x[kK], g[kK], v[kK];
    tile_load(p.x, c, D, cx.slice, live, x);
    tile_load(p.g, c, D, cx.slice, live, g);
    float lp = live ? p.lp[c] : 0.f;
    const float eps = tile_eps(p, c, live);
    const float heps = 0.5f * eps;
#pragma unroll 1
    for (int t = p.t0; t < p.t1; ++t) {
      const uint32_t* sk = p.sched + (size_t)(t - p.t_sched0) * p.sched_stride;
      const int r = tile_result_index(p, t);
      // ---- momentum draw (hmc.py:689-695) and first half kick (leapfrog_integrator.py:280-283)
      float s3[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        const int d = kK * cx.slice + j;
        const float m = (live && d < D) ? tile_momentum(p, sk, cg, d) : 0.f;
        s3[0] = fmaf(m, m, s3[0]);
        v[j] = m;
      }
      if (r >= 0 && p.tr.initial_momentum) tile_store(p.tr.initial_momentum, r, p.B, c, D, cx.slice, live, v);
#pragma unroll
      for (int j = 0; j < kK; ++j) v[j] = v[j] + heps * g[j];
      // ---- L leapfrogs; the gradient of all 128 chains is one tensor-core contraction
#pragma unroll 1
      for (int l = 0; l < p.L; ++l) {
#pragma unroll
        for (int j = 0; j < kK; ++j) x[j] = x[j] + eps * v[j];
        cx.stage_a(x);
        cx.contract();
        cx.load_d(g);
#pragma unroll
        for (int j = 0; j < kK; ++j) v[j] = v[j] + eps * g[j];
      }
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        v[j] = v[j] - heps * g[j];                 // final momentum
        s3[1] = fmaf(v[j], v[j], s3[1]);
        s3[2] = fmaf(x[j] - sh.loc[kK * cx.slice + j], g[j], s3[2]);
      }
      cx.reduce<3>(s3);
      const float lp1 = fmaf(0.5f, s3[2], tp.lognorm);
      const float corr = 0.5f * finite_or_neginf(s3[0] + (-s3[1]));      // hmc.py:862-875
      const float ratio = finite_or_neginf((lp1 + (-lp)) + corr);       // metropolis_hastings.py:204-215
      Key ka{sk[2 * p.n_parts], sk[2 * p.n_parts + 1]};
      const float u = uniform_from_bits(bits_at(ka, cg, (uint64_t)p.B_global, p.layout), 0.f, 1.f);
      const bool accept = logf(u) < ratio;                               // :221-227
      if (r >= 0) {
        const Trace& tr = p.tr;
        if (tr.proposed_state) tile_store(tr.proposed_state, r, p.B, c, D, cx.slice, live, x);
        if (tr.proposed_grads) tile_store(tr.proposed_grads, r, p.B, c, D, cx.slice, live, g);
        if (tr.final_momentum) tile_store(tr.final_momentum, r, p.B, c, D, cx.slice, live, v);
        if (live && cx.slice == 0) {
          const size_t o = (size_t)r * p.B + c;
          if (tr.proposed_target_log_prob) tr.proposed_target_log_prob[o] = lp1;
          if (tr.log_acceptance_correction) tr.log_acceptance_correction[o] = corr;
          if (tr.log_accept_ratio) tr.log_accept_ratio[o] = ratio;
          if (tr.is_accepted) tr.is_accepted[o] = accept ? 1 : 0;
        }
      }
      if (live && cx.slice == 0 && p.lar_last) p.lar_last[c] = ratio;
      if (accept) {
        lp = lp1;
        tile_store(p.x, 0, p.B, c, D, cx.slice, live, x);   // the chain-state arrays hold the accepted state
        tile_store(p.g, 0, p.B, c, D, cx.slice, live, g);
      } else {
        tile_load(p.x, c, D, cx.slice, live, x);
        tile_load(p.g, c, D, cx.slice, live, g);
      }
      if (r >= 0) {
        const Trace& tr = p.tr;
        if (tr.states) tile_store(tr.states, r, p.B, c, D, cx.slice, live, x);
        if (tr.grads) tile_store(tr.grads, r, p.B, c, D, cx.slice, live, g);
        if (live && cx.slice == 0) {
          if (tr.target_log_prob) tr.target_log_prob[(size_t)r * p.B + c] = lp;
          if (tr.step_size && c == 0 && p.step_kind == 0) tr.step_size[r] = p.step[0];
        }
      }
    }
    if (live && cx.slice == 0) {
      p.lp[c] = lp;
      if (p.leapfrog_total) p.leapfrog_total[c] += (unsigned long long)p.L * (unsigned long long)(p.t1 - p.t0);
    }
  }
  cx.finish();
}


// ---------------------------------------------------------------------------------------------
// tile_nuts_kernel: NoUTurnSampler.one_step (tfp/mcmc/nuts.py:321-946) for a tile of 128 chains run
// in LOCK-STEP, i.e. literally the reference's batched algorithm (shared doubling / leaf counters,
// per-chain masks), which makes every leapfrog of the tile one tensor-core contraction.
//   registers : moving trajectory end (x, m) and rho_subtree of my 26-dim slice, per-chain scalars (replicated x4)
//   TMEM      : g of the moving end (the contraction's accumulator) next to the MMA operands
//   smem      : the previous leaf's checkpoint (pb2_tile_nuts.cuh)
//   L2 scratch: other end, trajectory / subtree candidates, rho, the popcount-indexed checkpoint
//               stores -- per-thread segments in the 128-bit layout of pb2_tile.cuh
enum { kVOx = 0, kVOm, kVOg, kVCx, kVCg, kVBx, kVBg, kVRho, kVCk };   // checkpoints: kVCk + slot (m), + depth + slot (rho)

__global__ void __launch_bounds__(kThreads, 1)
tile_nuts_kernel(const ChainParams p, const DenseGaussianParams tp, float* __restrict__ scratch_all) {
  extern __shared__ __align__(128) unsigned char planes[];
  __shared__ Shared sh;
  __shared__ float lu[4][kM];   // log1p(-u) of the multinomial draws of 4 consecutive leaves
  Ctx cx;
  cx.init(&sh, planes, tp.P, tp.loc, tp.D);
  Prof pf;
  pf.init();
  const int D = tp.D;
  const int cl = cx.cl;
  const int nvec = kVCk + 2 * p.max_depth;
  // slice base of scratch vector v (the seg_* helpers add the chain)
  float* const scr_s = scratch_all + (size_t)blockIdx.x * nvec * kVS + (size_t)(kK * cx.slice) * kM;
  auto sv = [&](int v) -> float* { return scr_s + (size_t)v * kVS; };
  SubtreeArgs sa;
  sa.unrolled = p.unrolled; sa.layout = p.layout; sa.b_global = (uint64_t)p.B_global;
  sa.lognorm = tp.lognorm; sa.max_energy_diff = p.max_energy_diff;
  sa.lc = sh.loc + kK * cx.slice;
  sa.bx = sv(kVBx); sa.bg = sv(kVBg); sa.ck_m = sv(kVCk); sa.ck_r = sv(kVCk + p.max_depth);
  sa.ckl = reinterpret_cast<float*>(planes + 2 * kPlaneBytes) + (size_t)(kK * cx.slice) * kM;
  const int ntiles = (p.B + kM - 1) / kM;
  unsigned gt = 0;   // global leaf counter (rotates the "somebody continues" flags)
  for (int tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    const int c = tile_i * kM + cl;
    const bool live = c < p.B;
    const uint64_t cg = (uint64_t)p.chain_offset + (uint64_t)c;
    sa.cg = cg;
    float x[kK], m[kK], rho[kK];
    tile_load(p.x, c, D, cx.slice, live, x);
    {
      float g[kK];
      tile_load(p.g, c, D, cx.slice, live, g);
      cx.store_d(g);
    }
    float lp = live ? p.lp[c] : 0.f;
    unsigned long long nleap_total = 0;
#pragma unroll 1
    for (int t = p.t0; t < p.t1; ++t) {
      const float eps_abs = p.step_kind == 0 ? p.step[0] : (live ? p.step[c] : 0.f);
      const uint32_t* sk = p.sched + (size_t)(t - p.t_sched0) * p.sched_stride;
      const uint32_t* hdr = sk + 2 * p.n_parts;
      const uint32_t* ku = hdr + 6 * p.max_depth;
      const int r = tile_result_index(p, t);
      // ---- _start_trajectory_batched (nuts.py:512-539): momentum, H0; both ends, candidate, rho
      float s1[1] = {0.f};
      {
        float g[kK];
        cx.load_d(g);
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const int d = kK * cx.slice + j;
          const float mm = (live && d < D) ? tile_momentum(p, sk, cg, d) : 0.f;
          m[j] = mm;
          s1[0] = fmaf(mm, mm, s1[0]);
        }
        seg_st26(sv(kVOx), cl, x); seg_st26(sv(kVOm), cl, m); seg_st26(sv(kVOg), cl, g);
        seg_st26(sv(kVCx), cl, x); seg_st26(sv(kVCg), cl, g);
        seg_st26(sv(kVRho), cl, m);
      }
      cx.reduce<1>(s1);
      const float H0 = lp - 0.5f * s1[0];
      sa.H0 = H0;
      float slp = lp, olp = lp, clp = lp, cen = H0, cw = 0.f;
      float esum = 0.f;
      int nleap = 0;
      bool cont = live, notdiv = true, accepted = false, s_is_right = true;
      int any_cont = __syncthreads_or(cont ? 1 : 0);
#pragma unroll 1
      for (int it = 0; it < p.max_depth && any_cont; ++it) {
        // per-depth randoms of this chain (nuts.py:551-558, :622-625)
        Key kd{hdr[6 * it], hdr[6 * it + 1]}, kac{hdr[6 * it + 2], hdr[6 * it + 3]};
        const bool dir = (bits_at(kd, cg, (uint64_t)p.B_global, p.layout) & 1u) != 0;
        const float lacc = log1pf(-uniform_from_bits(bits_at(kac, cg, (uint64_t)p.B_global, p.layout), 0.f, 1.f));
        // registers / D must hold the end that is extended; _build_sub_tree init (nuts.py:713-791)
        {
          const bool sw = dir != s_is_right;
          float g[kK];
          cx.load_d(g);
          if (sw) {
            float o[kK];
            seg_ld26(sv(kVOx), cl, o); seg_st26(sv(kVOx), cl, x);
#pragma unroll
            for (int j = 0; j < kK; ++j) x[j] = o[j];
            seg_ld26(sv(kVOm), cl, o); seg_st26(sv(kVOm), cl, m);
#pragma unroll
            for (int j = 0; j < kK; ++j) m[j] = o[j];
            seg_ld26(sv(kVOg), cl, o); seg_st26(sv(kVOg), cl, g);
#pragma unroll
            for (int j = 0; j < kK; ++j) g[j] = o[j];
            const float a = slp; slp = olp; olp = a;
            s_is_right = dir;
          }
          if (__any_sync(0xffffffffu, sw)) cx.store_d(g);
          seg_st26(sa.bx, cl, x);
          seg_st26(sa.bg, cl, g);
        }
        sa.eps = dir ? eps_abs : -eps_abs;
        sa.nsteps = 1 << it;
        sa.kud = ku + 2 * (sa.nsteps - 1);
        SubtreeState st;
        st.slp = slp; st.c_prev = cont; st.nd = notdiv;
        nuts_subtree<false>(cx, sh, lu, gt, sa, x, m, rho, st, pf);
        slp = st.slp;
        const bool cont_f = st.c_prev;
        // _loop_tree_doubling tail (nuts.py:597-711)
        esum = st.esum_sub + esum;
        const float tw = cont_f ? st.bw : -INFINITY;
        const float wsum = log_add_exp(tw, cw);
        float thr = tw - cw;
        thr = isnan(thr) ? 0.f : thr;
        const bool swap = (lacc <= thr) && cont_f;
        cw = wsum;
        if (swap) {
          float o[kK];
          seg_ld26(sa.bx, cl, o); seg_st26(sv(kVCx), cl, o);
          seg_ld26(sa.bg, cl, o); seg_st26(sv(kVCg), cl, o);
          clp = st.blp; cen = st.ben;
        }
        float s2[2] = {0.f, 0.f};
        {
          float rh[kK], om[kK];
          seg_ld26(sv(kVRho), cl, rh);
          seg_ld26(sv(kVOm), cl, om);
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            const float rr = rh[j] + rho[j];
            rh[j] = rr;
            s2[0] = fmaf(rr, m[j], s2[0]);
            s2[1] = fmaf(rr, om[j], s2[1]);
          }
          seg_st26(sv(kVRho), cl, rh);
        }
        cx.reduce<2>(s2);
        nleap += st.n;
        accepted = accepted || swap;
        notdiv = st.nd;
        cont = cont_f && (s2[0] >= 0.f) && (s2[1] >= 0.f);
        any_cont = __syncthreads_or(cont ? 1 : 0);       // nuts.py:404-407
      }
      // ---- results (nuts.py:424-445); the next state is the trajectory candidate
      float g[kK];
      seg_ld26(sv(kVCx), cl, x);
      seg_ld26(sv(kVCg), cl, g);
      cx.store_d(g);
      lp = clp;
      const int leap = nleap * p.unrolled;
      nleap_total += (unsigned long long)leap;
      const float lar = logf(esum / (float)nleap);
      if (live && cx.slice == 0 && p.lar_last) p.lar_last[c] = lar;
      if (r >= 0) {
        const Trace& tr = p.tr;
        if (tr.states) tile_store(tr.states, r, p.B, c, D, cx.slice, live, x);
        if (tr.grads) tile_store(tr.grads, r, p.B, c, D, cx.slice, live, g);
        if (live && cx.slice == 0) {
          const size_t o = (size_t)r * p.B + c;
          if (tr.target_log_prob) tr.target_log_prob[o] = lp;
          if (tr.log_accept_ratio) tr.log_accept_ratio[o] = lar;
          if (tr.is_accepted) tr.is_accepted[o] = accepted ? 1 : 0;
          if (tr.leapfrogs_taken) tr.leapfrogs_taken[o] = leap;
          if (tr.has_divergence) tr.has_divergence[o] = notdiv ? 0 : 1;
          if (tr.reach_max_depth) tr.reach_max_depth[o] = cont ? 1 : 0;
          if (tr.energy) tr.energy[o] = cen;
          if (tr.step_size && c == 0 && p.step_kind == 0) tr.step_size[r] = p.step[0];
        }
      }
    }
    {
      float g[kK];
      cx.load_d(g);
      tile_store(p.x, 0, p.B, c, D, cx.slice, live, x);
      tile_store(p.g, 0, p.B, c, D, cx.slice, live, g);
    }
    if (live && cx.slice == 0) {
      p.lp[c] = lp;
      if (p.leapfrog_total) p.leapfrog_total[c] += nleap_total;
    }
  }
  cx.finish();
}

}  // namespace pb2
#include "pb2_tile_sched.cuh"
namespace pb2 {

bool tile_path_supported(const pb2_ctx* ctx, const pb2_target* tgt, int mode, const ChainParams& p) {
  if (ctx->dense_variant == 1) return false;                       // PB2_DENSE_VARIANT=1: force warp-per-chain
  if (tgt->kind != PB2_TARGET_DENSE_GAUSSIAN) return false;
  if (tgt->dim <= 32 || tgt->dim > 100) return false;
  if (mode != kModeHMC && mode != kModeNUTS) return false;
  if (p.step_kind == 1) return false;                              // per-dimension step sizes: warp kernels
  return p.B >= 2 * kM;
}

#ifdef PB2_TILE_PROF
static void dump_tile_prof(pb2_ctx* ctx) {
  unsigned long long h[2][16];
  cudaStreamSynchronize(ctx->stream);
  cudaMemcpyFromSymbol(h, tile::g_tile_prof, sizeof(h));
  static const char* nm[7] = {"head", "kick+stage", "contract", "post", "reduce4", "extra checks", "scalars+take"};
  for (int w = 0; w < 2; ++w) {
    fprintf(stderr, "[tileprof t%d] leaves %llu:", w ? 511 : 0, h[w][15]);
    for (int k = 0; k < 7; ++k) fprintf(stderr, " %s %.0f", nm[k], h[w][15] ? (double)h[w][k] / h[w][15] : 0.0);
    fprintf(stderr, "\n");
  }
  unsigned long long z[2][16] = {};
  cudaMemcpyToSymbol(tile::g_tile_prof, z, sizeof(z));
}
#else
static void dump_tile_prof(pb2_ctx*) {}
#endif

int launch_tile_chain(pb2_ctx* ctx, const pb2_target* tgt, int mode, ChainParams& p) {
  DenseGaussianParams tp{tgt->d_a, tgt->d_b, tgt->scalar, tgt->dim};
  size_t smem = 2 * (size_t)kPlaneBytes;
  const int ntiles = (p.B + kM - 1) / kM;
  const int grid = std::min(ntiles, ctx->num_sms);
  if (mode == kModeHMC) {
    if (int rc = check_cuda(ctx, cudaFuncSetAttribute(tile_hmc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)smem), "cudaFuncSetAttribute(tile_hmc)"))
      return rc;
    tile_hmc_kernel<<<grid, kThreads, smem, ctx->stream>>>(p, tp);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "tile_hmc_kernel");
  }
  smem += 2 * kVS * sizeof(float);   // NUTS: + the previous leaf's checkpoint (momentum, rho)
  // fused multi-transition NUTS runs: chains re-grouped every 32 leaves (pb2_tile_sched.cuh)
  if (mode == kModeNUTS && ctx->dense_variant != 3 && p.lar_last == nullptr && p.t1 - p.t0 >= 2 &&
      p.max_depth > kS0) {
    const int sgrid = getenv("PB2_SCHED_GRID") ? atoi(getenv("PB2_SCHED_GRID")) : ctx->num_sms;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    SchedParams sp;
    sp.nrv = kRHi + 2 * (p.max_depth - kS0);
    const size_t scr_bytes = up((size_t)sgrid * (2 + 2 * kS0) * kVS * sizeof(float));
    const size_t vec_bytes = up((size_t)p.B * sp.nrv * kRecStride * sizeof(float));
    const size_t scal_bytes = up((size_t)p.B * kRecScal * sizeof(float));
    const size_t queue_bytes = up((size_t)2 * p.B * sizeof(int));
    const size_t need = scr_bytes + vec_bytes + scal_bytes + queue_bytes + kQWords * sizeof(unsigned long long);
    if (need > ctx->ckpt_bytes) {
      if (ctx->d_ckpt) cudaFree(ctx->d_ckpt);
      ctx->d_ckpt = nullptr;
      ctx->ckpt_bytes = 0;
      if (int rc = check_cuda(ctx, cudaMalloc(&ctx->d_ckpt, need), "cudaMalloc(tile scheduler)")) return rc;
      ctx->ckpt_bytes = need;
    }
    unsigned char* base = reinterpret_cast<unsigned char*>(ctx->d_ckpt);
    sp.rec_vec = reinterpret_cast<float*>(base + scr_bytes);
    sp.rec_scal = reinterpret_cast<float*>(base + scr_bytes + vec_bytes);
    sp.queue = reinterpret_cast<int*>(base + scr_bytes + vec_bytes + scal_bytes);
    sp.qctl = reinterpret_cast<unsigned long long*>(base + scr_bytes + vec_bytes + scal_bytes + queue_bytes);
    sp.patience = getenv("PB2_SCHED_PATIENCE") ? atoi(getenv("PB2_SCHED_PATIENCE")) : 50;
    sp.stats = nullptr;
    static unsigned long long* d_stats = nullptr;
    if (getenv("PB2_SCHED_STATS")) {
      if (!d_stats) cudaMalloc(&d_stats, 32 * sizeof(unsigned long long));
      cudaMemsetAsync(d_stats, 0, 32 * sizeof(unsigned long long), ctx->stream);
      sp.stats = d_stats;
    }
    tile_sched_init_kernel<<<(p.B + 255) / 256, 256, 0, ctx->stream>>>(sp, p.B, p.t0);
    if (int rc = check_cuda(ctx, cudaFuncSetAttribute(tile_nuts_sched_kernel,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(tile_nuts_sched)"))
      return rc;
    tile_nuts_sched_kernel<<<sgrid, kThreads, smem, ctx->stream>>>(p, tp, sp, ctx->d_ckpt);
    ctx->launches += 2;
    dump_tile_prof(ctx);
    if (sp.stats) {
      unsigned long long h[32];
      cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      fprintf(stderr, "[sched] tasks %llu chains %llu (fill %.1f) ticks %llu idle-polls %llu | per class tasks:", h[0], h[1],
              h[0] ? (double)h[1] / h[0] : 0.0, h[2], h[3]);
      for (int k = 0; k < 8; ++k) fprintf(stderr, " %llu(%.0f)", h[8 + k], h[8 + k] ? (double)h[20 + k] / h[8 + k] : 0.0);
      fprintf(stderr, "\n[sched] Mcycles summed over CTAs: claim %.1f idle %.1f load %.1f run %.1f publish %.1f (run = %.0f cycles per tile leaf)\n",
              h[26] * 1e-6, h[27] * 1e-6, h[28] * 1e-6, h[29] * 1e-6, h[30] * 1e-6, h[2] ? (double)h[29] / h[2] : 0.0);
    }
    return check_cuda(ctx, cudaGetLastError(), "tile_nuts_sched_kernel");
  }
  if (mode == kModeNUTS) {
    const size_t per_cta = (size_t)(kVCk + 2 * p.max_depth) * kKP * kM * sizeof(float);
    const size_t need = per_cta * grid;
    if (need > ctx->ckpt_bytes) {
      if (ctx->d_ckpt) cudaFree(ctx->d_ckpt);
      ctx->d_ckpt = nullptr;
      ctx->ckpt_bytes = 0;
      if (int rc = check_cuda(ctx, cudaMalloc(&ctx->d_ckpt, need), "cudaMalloc(tile scratch)")) return rc;
      ctx->ckpt_bytes = need;
    }
    if (int rc = check_cuda(ctx, cudaFuncSetAttribute(tile_nuts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)smem), "cudaFuncSetAttribute(tile_nuts)"))
      return rc;
    tile_nuts_kernel<<<grid, kThreads, smem, ctx->stream>>>(p, tp, ctx->d_ckpt);
    ctx->launches += 1;
    dump_tile_prof(ctx);
    return check_cuda(ctx, cudaGetLastError(), "tile_nuts_kernel");
  }
  return set_error(ctx, PB2_ERR_UNSUPPORTED, "tile path: unsupported mode");
}

}  // namespace pb2
