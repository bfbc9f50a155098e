// Bayesian logistic regression log-prob + gradient for ALL chains at once on the 5th-gen tensor cores:
//     z = Theta X~^T                                  [chains x N]     GEMM 1 (K = D)
//     lp = sum_d N(theta_d; 0, 1) + sum_n [y_n z_n - softplus(z_n)]     (gym logistic_regression.py:88-103,
//     g  = -theta + (y - sigmoid(z)) X~               [chains x D]     GEMM 2 (K = N)      bernoulli.py:119-135)
// fused like an attention block: the [128 x N] logits never leave the SM.  One CTA = one tile of 128 chains; the N
// rows of X~ are processed in chunks of 64 (the comments below describe one chunk):
//     A1 = Theta hi/lo (TMEM, 32 columns each)                B1 = X~ chunk [n = 128 rows][k = 32] K-major hi/lo (smem)
//     D1 = z chunk [128 x 128] (TMEM)  -> epilogue: lp += y z - softplus z ; r = y - sigmoid z -> A2 = r hi/lo (TMEM)
//     D2 += A2 [128 x K = 128 rows] . B2,  B2 = X~ chunk [n = 32 dims][k = 128 rows] K-major hi/lo (smem)
// 3xTF32 split on both GEMMs (FP32-accurate): (3 x 4 + 3 x 16) MMAs per chunk.  TMEM: 64 + 128 + 256 + 32 = 480 columns.
// This is the primitive (SURVEY 8a row T3 on tcgen05); the transition kernels still use the FP32 warp-per-chain
// gradient (pb2_targets.cuh LogisticT).
#include <algorithm>
#include "pb2_tile.cuh"

namespace pb2 {
using tile::make_kmajor_desc;
using tile::mbar_wait;
using tile::smem_u32;
using tile::tmem_ld;
using tile::tmem_st;
using tile::tmem_wait_ld;

namespace ltc {
constexpr int kM = 128;        // chains per tile
constexpr int kKD = 32;        // padded D (K of GEMM 1, N of GEMM 2)
constexpr int kR = 64;         // rows of X~ per chunk (N of GEMM 1, K of GEMM 2)
constexpr int kWorkers = 512;  // 16 worker warps: thread = (chain row, slice of 16 of the chunk's 64 rows of X~)
constexpr int kThreads = kWorkers + 32;   // + warp 16: issues the contractions
constexpr int kRing = 4;       // chunk operand buffers: GEMM 2 of c - 1 and c, GEMM 1 of c + 1, async copy of c + 2
// TMEM columns: theta hi/lo | 2 x z chunk | 2 x r hi | 2 x r lo | 2 x g chunk
constexpr int kColA1hi = 0, kColA1lo = 32, kColD1 = 64, kColA2hi = 192, kColA2lo = 320, kColD2 = 448;
constexpr int kB1Plane = kR * kKD * 4;   // 8 KB: [n = 64 rows][k = 32 dims]
constexpr int kB2Plane = kKD * kR * 4;   // 8 KB: [n = 32 dims][k = 64 rows]
constexpr int kChunkBytes = 2 * kB1Plane + 2 * kB2Plane;   // hi/lo of both layouts = 32 KB

// byte offset of element (n, k) of a [NP x KP] K-major no-swizzle operand made of 8 x 16 B core matrices, K-chunk
// major: LBO = (NP/8)*128 B between the two K core matrices of one MMA, SBO = 128 B between row groups
template <int NP>
__host__ __device__ inline int plane_offset(int n, int k) {
  return ((k >> 2) * (NP / 8) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 3) * 4;
}

__device__ __forceinline__ uint32_t tf32_round(float v) { return (__float_as_uint(v) + 0x1000u) & 0xffffe000u; }

// ---- one-time: X~ [N, D] -> per chunk the four planes (B1 hi, B1 lo, B2 hi, B2 lo) in the canonical layouts
__global__ void logistic_tc_prepare_kernel(const float* __restrict__ X, int N, int D, unsigned char* __restrict__ out,
                                           int nchunks) {
  const int chunk = blockIdx.x;
  unsigned char* o = out + (size_t)chunk * kChunkBytes;
  for (int i = threadIdx.x; i < kR * kKD; i += blockDim.x) {
    const int r = i / kKD, d = i - r * kKD;   // row of the chunk, dim
    const int n = chunk * kR + r;
    const float v = (n < N && d < D) ? X[(size_t)n * D + d] : 0.f;
    const uint32_t hi = tf32_round(v);
    const uint32_t lo = tf32_round(v - __uint_as_float(hi));
    const int o1 = plane_offset<kR>(r, d);     // B1[n = r][k = d]
    const int o2 = plane_offset<kKD>(d, r);    // B2[n = d][k = r]
    *reinterpret_cast<uint32_t*>(o + o1) = hi;
    *reinterpret_cast<uint32_t*>(o + kB1Plane + o1) = lo;
    *reinterpret_cast<uint32_t*>(o + 2 * kB1Plane + o2) = hi;
    *reinterpret_cast<uint32_t*>(o + 2 * kB1Plane + kB2Plane + o2) = lo;
  }
}

struct Smem {
  unsigned long long g1_done[2], g2_done[2];   // mbarriers: the contraction into D1[b] / D2[b] has completed
  uint32_t tmem_base;
  int quit;
  float y[kRing][kR];
  float valid[kRing][kR];
  float red[4][kM];
};

// Software pipeline (per 128-chain tile, chunk c of 64 rows of X~):
//   workers : start the async copy of chunk c+2's operands -> wait z(c) -> read it (D1[c%2] is free) -> signal A -> sigmoid /
//             softplus of the 16 logits of my slice -> r hi/lo into A2[c%2] -> signal B
//   issuer  : on A: GEMM 1 of chunk c+2 into D1[c%2];  on B: GEMM 2 of chunk c into D2[c%2]
// so both contractions run in the shadow of the workers' transcendental work.  UTCHMMA issue is back-pressured by
// the tensor pipe, which is why it lives in a warp of its own.  Signals A / B are named barriers 2 / 3 on which the
// workers only arrive.
__global__ void __launch_bounds__(kThreads, 1)
logistic_tc_kernel(const float* __restrict__ Theta, int B, int D, int N, const unsigned char* __restrict__ planes_g,
                   const float* __restrict__ labels, int nchunks, float* __restrict__ out_lp, float* __restrict__ out_g) {
  extern __shared__ __align__(128) unsigned char ring[];   // kRing x (B1 hi, B1 lo, B2 hi, B2 lo)
  __shared__ Smem sh;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.g1_done[b])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.g2_done[b])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
    sh.quit = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = sh.tmem_base;
  const int ntiles = (B + kM - 1) / kM;
  const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  if (warp == kWorkers / 32) {
    // ------------------------------------------------------------------ the issuing warp
    const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kR >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kKD >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    uint32_t leader;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
    auto gemm1 = [&](int c) {   // z chunk = theta . X~chunk^T : 3 passes x 4 K-steps, M128 N64 K8
      if (!leader) return;
      const uint32_t base = smem_u32(ring + (size_t)(c % kRing) * kChunkBytes);
      const uint64_t bhi = make_kmajor_desc(base, (kR / 8) * 128, 128);
      const uint64_t blo = make_kmajor_desc(base + kB1Plane, (kR / 8) * 128, 128);
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
        for (int j = 0; j < kKD / 8; ++j) {
          const uint32_t a = tmem + (pass == 1 ? kColA1lo : kColA1hi) + 8 * j;
          const uint64_t bd = (pass == 2 ? blo : bhi) + (uint64_t)(j * ((2u * (kR / 8) * 128u) >> 4));
          const uint32_t acc = (pass == 0 && j == 0) ? 0u : 1u;
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + kColD1 + kR * (c & 1)),
              "r"(a), "l"(bd), "r"(idesc1), "r"(acc)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&sh.g1_done[c & 1]))
                   : "memory");
    };
    auto gemm2 = [&](int c) {   // g chunk = r . X~chunk : 3 passes x 8 K-steps, M128 N32 K8
      if (!leader) return;
      const uint32_t base = smem_u32(ring + (size_t)(c % kRing) * kChunkBytes) + 2 * kB1Plane;
      const uint64_t bhi = make_kmajor_desc(base, (kKD / 8) * 128, 128);
      const uint64_t blo = make_kmajor_desc(base + kB2Plane, (kKD / 8) * 128, 128);
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
        for (int j = 0; j < kR / 8; ++j) {
          const uint32_t a = tmem + (pass == 1 ? kColA2lo : kColA2hi) + kR * (c & 1) + 8 * j;
          const uint64_t bd = (pass == 2 ? blo : bhi) + (uint64_t)(j * ((2u * (kKD / 8) * 128u) >> 4));
          const uint32_t acc = (pass == 0 && j == 0) ? 0u : 1u;   // per-chunk product; chunks are summed in registers
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + kColD2 + kKD * (c & 1)),
              "r"(a), "l"(bd), "r"(idesc2), "r"(acc)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&sh.g2_done[c & 1]))
                   : "memory");
    };
    for (int tl = 0; tl < my_tiles; ++tl) {
      asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");   // theta staged, chunks 0 and 1 copied
      asm volatile("tcgen05.fence::after_thread_sync;");
      gemm1(0);
      if (nchunks > 1) gemm1(1);
      for (int c = 0; c < nchunks; ++c) {
        asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");   // D1[c%2] read, chunk c+2 copied
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (c + 2 < nchunks) gemm1(c + 2);
        asm volatile("bar.sync 3, %0;" ::"n"(kThreads) : "memory");   // A2[c%2] written
        asm volatile("tcgen05.fence::after_thread_sync;");
        gemm2(c);
      }
    }
  } else {
    // ------------------------------------------------------------------ the workers
    const int row = 32 * (warp & 3) + (tid & 31);   // chain of the tile = TMEM lane
    const int slice = warp >> 2;                    // 16 of the chunk's 64 rows of X~ / 8 of the 32 dims
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    unsigned n1[2] = {0u, 0u}, n2[2] = {0u, 0u};    // completed waits per mbarrier (-> its phase parity)
    auto stage = [&](int c) {   // chunk c's operand planes and labels -> ring slot c % kRing (cp.async, 16 B each)
      const unsigned char* src = planes_g + (size_t)c * kChunkBytes;
      const uint32_t dst = smem_u32(ring + (size_t)(c % kRing) * kChunkBytes);
      for (int i = tid; i < kChunkBytes / 16; i += kWorkers)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * i), "l"(src + 16 * (size_t)i) : "memory");
      if (tid < kR) {
        const int n = c * kR + tid;
        sh.y[c % kRing][tid] = n < N ? labels[n] : 0.f;
        sh.valid[c % kRing][tid] = n < N ? 1.f : 0.f;
      }
    };
    auto signal = [&](int bar) {   // workers only ARRIVE (after making their copies / generic / TMEM writes visible)
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      if (bar == 2) asm volatile("bar.arrive 2, %0;" ::"n"(kThreads) : "memory");
      else asm volatile("bar.arrive 3, %0;" ::"n"(kThreads) : "memory");
    };
    for (int tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
      const int c0 = tile_i * kM + row;
      const bool live = c0 < B;
      // ---- A1 = theta hi/lo: my 8 dims (slice) of my chain
      float th[8];
      float prior = 0.f;
      {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int d = 8 * slice + j;
          th[j] = (live && d < D) ? Theta[(size_t)c0 * D + d] : 0.f;
          if (d < D) prior += -0.5f * th[j] * th[j] - 0.9189385332046727f;
          hi[j] = tf32_round(th[j]);
          lo[j] = tf32_round(th[j] - __uint_as_float(hi[j]));
        }
        tmem_st<8>(lane_addr + kColA1hi + 8 * slice, hi);
        tmem_st<8>(lane_addr + kColA1lo + 8 * slice, lo);
      }
      stage(0);
      if (nchunks > 1) stage(1);
      signal(2);
      float ll = 0.f;   // my share of sum_n [y z - softplus z]
      float gacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      auto take_g = [&](int c) {   // g chunk of chunk c: wait, add (FP32 adds keep the tensor core's sums short)
        const int b = c & 1;
        mbar_wait(smem_u32(&sh.g2_done[b]), n2[b] & 1u);
        n2[b]++;
        asm volatile("tcgen05.fence::after_thread_sync;");
        uint32_t gq[8];
        tmem_ld<8>(lane_addr + kColD2 + kKD * b + 8 * slice, gq);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) gacc[j] += __uint_as_float(gq[j]);
      };
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        if (c + 2 < nchunks) stage(c + 2);   // slot (c+2) % 4 was last read by chunk c-2's contractions (waited for)
        mbar_wait(smem_u32(&sh.g1_done[b]), n1[b] & 1u);
        n1[b]++;
        asm volatile("tcgen05.fence::after_thread_sync;");
        uint32_t zq[16];
        tmem_ld<16>(lane_addr + kColD1 + kR * b + 16 * slice, zq);
        tmem_wait_ld();
        // chunk c-1's contraction 2 is the last reader of A2[(c-1)%2] and the writer of D2[(c-1)%2]
        if (c >= 1) take_g(c - 1);
        signal(2);
        // ---- my 16 logits -> log-likelihood terms, r = y - sigmoid(z) -> A2[b] hi/lo
        uint32_t hi[16], lo[16];
        const float* yy = sh.y[c % kRing] + 16 * slice;
        const float* vv = sh.valid[c % kRing] + 16 * slice;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float z = __uint_as_float(zq[j]);
          // softplus(z) = max(z, 0) + log1p(exp(-|z|)); sigmoid(z) from the same exponential
          const float e = __expf(-fabsf(z));                 // same fast forms as the FP32 kernel (pb2_targets.cuh)
          const float sp = fmaxf(z, 0.f) + __logf(1.0f + e);
          const float inv = __fdividef(1.0f, 1.0f + e);
          const float sg = z >= 0.f ? inv : e * inv;
          ll += vv[j] * (yy[j] * z - sp);
          const float r = vv[j] * (yy[j] - sg);
          hi[j] = tf32_round(r);
          lo[j] = tf32_round(r - __uint_as_float(hi[j]));
        }
        tmem_st<16>(lane_addr + kColA2hi + kR * b + 16 * slice, hi);
        tmem_st<16>(lane_addr + kColA2lo + kR * b + 16 * slice, lo);
        signal(3);
      }
      take_g(nchunks - 1);
      // ---- g = sum of chunks - theta ; lp = prior + log-likelihood, summed over the 4 slices in a fixed order
      if (live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int d = 8 * slice + j;
          if (d < D) out_g[(size_t)c0 * D + d] = gacc[j] - th[j];
        }
      }
      sh.red[slice][row] = ll + prior;
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
      if (live && slice == 0) out_lp[c0] = ((sh.red[0][row] + sh.red[1][row]) + sh.red[2][row]) + sh.red[3][row];
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
}  // namespace ltc

int launch_logistic_tc(pb2_ctx* ctx, pb2_target* tgt, int B, const float* d_x, float* d_lp, float* d_g) {
  using namespace ltc;
  const int D = tgt->dim, N = tgt->n_rows;
  if (D > kKD) return set_error(ctx, PB2_ERR_UNSUPPORTED, "pb2_logistic_logp_grad_tc: D <= 32 only");
  const int nchunks = (N + kR - 1) / kR;
  const size_t need = (size_t)nchunks * kChunkBytes;
  if (!tgt->d_tc) {   // the operand planes of this target, built once
    if (int rc = check_cuda(ctx, cudaMalloc(&tgt->d_tc, need), "cudaMalloc(logistic tc planes)")) return rc;
    tgt->tc_bytes = need;
    logistic_tc_prepare_kernel<<<nchunks, 256, 0, ctx->stream>>>(tgt->d_a, N, D, tgt->d_tc, nchunks);
    ctx->launches += 1;
  }
  const size_t smem = (size_t)kRing * kChunkBytes;
  if (int rc = check_cuda(ctx, cudaFuncSetAttribute(logistic_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(logistic_tc)"))
    return rc;
  const int grid = std::min((B + kM - 1) / kM, ctx->num_sms);
  logistic_tc_kernel<<<grid, kThreads, smem, ctx->stream>>>(d_x, B, D, N, tgt->d_tc, tgt->d_b, nchunks, d_lp, d_g);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "logistic_tc_kernel");
}

}  // namespace pb2

extern "C" int pb2_logistic_logp_grad_tc(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_x, float* d_logp,
                                         float* d_grad) {
  if (!ctx || !tgt || B < 1 || !d_x || !d_logp || !d_grad)
    return pb2::set_error(ctx, PB2_ERR_INVALID, "pb2_logistic_logp_grad_tc: bad argument");
  if (tgt->kind != PB2_TARGET_LOGISTIC)
    return pb2::set_error(ctx, PB2_ERR_INVALID, "pb2_logistic_logp_grad_tc: target must be a logistic regression");
  return pb2::launch_logistic_tc(ctx, const_cast<pb2_target*>(tgt), B, d_x, d_logp, d_grad);
}
