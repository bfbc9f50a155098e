// 128-chain TILE transition kernels for the dense-Gaussian target (tcgen05 path).
//   tile_hmc_kernel : HamiltonianMonteCarlo = MetropolisHastings(UncalibratedHMC)
//                     (tfp/mcmc/hmc.py:661-729,780-875; metropolis_hastings.py:181-254)
// All chains of a tile advance in lock-step; each leapfrog's gradient is one 3xTF32 tensor-core
// contraction (pb2_tile.cuh).  Same seeds, same counters, same decisions as the warp-per-chain
// kernels and the oracle (the uint32 streams are identical; floats agree to rounding).
#include <cstdio>
#include <cstdlib>
#include "pb2_tile.cuh"

namespace pb2 {
using namespace tile;

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
  cx.init(&sh, planes, tp.P, tp.loc, tp.D, tp.scale);
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


}  // namespace pb2

namespace pb2 {

bool tile_path_supported(const pb2_ctx* ctx, const pb2_target* tgt, int mode, const ChainParams& p) {
  if (ctx->dense_variant == 1) return false;                       // PB2_DENSE_VARIANT=1: force warp-per-chain
  if (tgt->kind != PB2_TARGET_DENSE_GAUSSIAN) return false;
  if (tgt->dim <= 32 || tgt->dim > 100) return false;
  if (mode != kModeHMC && mode != kModeNUTS) return false;
  if (p.step_kind == 1) return false;                              // per-dimension step sizes: warp kernels
  if (p.bij_kind) return false;                                    // event-space bijectors: warp kernels
  // the choice depends on the chains of the WHOLE job, so that every shard of a chain-sharded run takes the same path
  return std::max(p.B, p.B_global) >= 2 * kM;
}

int launch_tile_chain(pb2_ctx* ctx, const pb2_target* tgt, int mode, ChainParams& p) {
  DenseGaussianParams tp{tgt->d_a, tgt->d_b, tgt->scalar, tgt->dim};
  tp.scale = p.scale;
  const size_t smem = 2 * (size_t)kPlaneBytes;
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
  if (mode == kModeNUTS) return launch_tile_nuts(ctx, tgt, p);
  return set_error(ctx, PB2_ERR_UNSUPPORTED, "tile path: unsupported mode");
}

}  // namespace pb2
