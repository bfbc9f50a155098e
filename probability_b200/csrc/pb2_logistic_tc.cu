// Bayesian logistic regression log-prob + gradient for ALL chains at once on the 5th-gen tensor cores:
//     z = Theta X~^T                                  [chains x N]     GEMM 1 (K = D)
//     lp = sum_d N(theta_d; 0, 1) + sum_n [y_n z_n - softplus(z_n)]     (gym logistic_regression.py:88-103,
//     g  = -theta + (y - sigmoid(z)) X~               [chains x D]     GEMM 2 (K = N)      bernoulli.py:119-135)
// fused like an attention block: the [128 x N] logits never leave the SM.  One CTA = one tile of 128 chains; the N
// rows of X~ are processed in chunks of 128:
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
constexpr int kR = 128;        // rows of X~ per chunk (N of GEMM 1, K of GEMM 2)
constexpr int kThreads = 512;
constexpr int kColA1hi = 0, kColA1lo = 32, kColD1 = 64, kColA2hi = 192, kColA2lo = 320, kColD2 = 448;
constexpr int kB1Plane = kR * kKD * 4;   // 16 KB: [n = 128][k = 32]
constexpr int kB2Plane = kKD * kR * 4;   // 16 KB: [n = 32][k = 128]
constexpr int kChunkBytes = 2 * kB1Plane + 2 * kB2Plane;   // hi/lo of both layouts

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
  unsigned long long mbar;
  uint32_t tmem_base;
  float y[kR];
  float valid[kR];
  float red[4][kM];
};

__global__ void __launch_bounds__(kThreads, 1)
logistic_tc_kernel(const float* __restrict__ Theta, int B, int D, int N, const unsigned char* __restrict__ planes_g,
                   const float* __restrict__ labels, int nchunks, float* __restrict__ out_lp, float* __restrict__ out_g) {
  extern __shared__ __align__(128) unsigned char planes[];   // one chunk: B1 hi, B1 lo, B2 hi, B2 lo
  __shared__ Smem sh;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = 32 * (warp & 3) + (tid & 31);   // chain of the tile = TMEM lane
  const int slice = warp >> 2;                    // 32 of the chunk's 128 rows of X~ / 8 of the 32 dims
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = sh.tmem_base;
  const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint64_t b1hi = make_kmajor_desc(smem_u32(planes), (kR / 8) * 128, 128);
  const uint64_t b1lo = make_kmajor_desc(smem_u32(planes + kB1Plane), (kR / 8) * 128, 128);
  const uint64_t b2hi = make_kmajor_desc(smem_u32(planes + 2 * kB1Plane), (kKD / 8) * 128, 128);
  const uint64_t b2lo = make_kmajor_desc(smem_u32(planes + 2 * kB1Plane + kB2Plane), (kKD / 8) * 128, 128);
  // instruction descriptors: D = f32, A = B = tf32, K-major, N >> 3, M >> 4
  const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kR >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
  const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kKD >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
  uint32_t phase = 0;

  const int ntiles = (B + kM - 1) / kM;
  for (int tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    const int c = tile_i * kM + row;
    const bool live = c < B;
    // ---- A1 = theta hi/lo: my 8 dims (slice) of my chain
    float th[8];
    float prior = 0.f;
    {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int d = 8 * slice + j;
        th[j] = (live && d < D) ? Theta[(size_t)c * D + d] : 0.f;
        if (d < D) prior += -0.5f * th[j] * th[j] - 0.9189385332046727f;
        hi[j] = tf32_round(th[j]);
        lo[j] = tf32_round(th[j] - __uint_as_float(hi[j]));
      }
      tmem_st<8>(lane_addr + kColA1hi + 8 * slice, hi);
      tmem_st<8>(lane_addr + kColA1lo + 8 * slice, lo);
    }
    float ll = 0.f;   // my share of sum_n [y z - softplus z]
    float gacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int ch = 0; ch < nchunks; ++ch) {
      // ---- stage the chunk's operand planes and labels (plain 128-bit copies; the previous chunk's MMAs are done)
      {
        const uint4* src = reinterpret_cast<const uint4*>(planes_g + (size_t)ch * kChunkBytes);
        uint4* dst = reinterpret_cast<uint4*>(planes);
        for (int i = tid; i < kChunkBytes / 16; i += kThreads) dst[i] = src[i];
        if (tid < kR) {
          const int n = ch * kR + tid;
          sh.y[tid] = n < N ? labels[n] : 0.f;
          sh.valid[tid] = n < N ? 1.f : 0.f;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();
      // ---- GEMM 1: z chunk = theta . X~chunk^T   (3 passes x 4 K-steps, M128 N128 K8)
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
          for (int j = 0; j < kKD / 8; ++j) {
            const uint32_t a = tmem + (pass == 1 ? kColA1lo : kColA1hi) + 8 * j;
            const uint64_t bd = (pass == 2 ? b1lo : b1hi) + (uint64_t)(j * ((2u * (kR / 8) * 128u) >> 4));
            const uint32_t acc = (pass == 0 && j == 0) ? 0u : 1u;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + kColD1),
                "r"(a), "l"(bd), "r"(idesc1), "r"(acc)
                : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sh.mbar))
                     : "memory");
        mbar_wait(smem_u32(&sh.mbar), phase);
        asm volatile("tcgen05.fence::before_thread_sync;");
      }
      phase ^= 1;
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;");
      // ---- epilogue 1: my 32 logits -> log-likelihood terms, r = y - sigmoid(z) -> A2 hi/lo
      {
        uint32_t z0[16], z1[16];
        tmem_ld<16>(lane_addr + kColD1 + 32 * slice, z0);
        tmem_ld<16>(lane_addr + kColD1 + 32 * slice + 16, z1);
        tmem_wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float z = __uint_as_float(half ? z1[j] : z0[j]);
            const int rr = 32 * slice + 16 * half + j;
            const float y = sh.y[rr], v = sh.valid[rr];
            // softplus(z) = max(z, 0) + log1p(exp(-|z|)); sigmoid(z) from the same exponential
            const float e = expf(-fabsf(z));
            const float sp = fmaxf(z, 0.f) + log1pf(e);
            const float inv = 1.f / (1.f + e);
            const float sg = z >= 0.f ? inv : e * inv;
            ll += v * (y * z - sp);
            const float r = v * (y - sg);
            hi[j] = tf32_round(r);
            lo[j] = tf32_round(r - __uint_as_float(hi[j]));
          }
          tmem_st<16>(lane_addr + kColA2hi + 32 * slice + 16 * half, hi);
          tmem_st<16>(lane_addr + kColA2lo + 32 * slice + 16 * half, lo);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();
      // ---- GEMM 2: g += r . X~chunk   (3 passes x 16 K-steps, M128 N32 K8)
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
          for (int j = 0; j < kR / 8; ++j) {
            const uint32_t a = tmem + (pass == 1 ? kColA2lo : kColA2hi) + 8 * j;
            const uint64_t bd = (pass == 2 ? b2lo : b2hi) + (uint64_t)(j * ((2u * (kKD / 8) * 128u) >> 4));
            const uint32_t acc = (pass == 0 && j == 0) ? 0u : 1u;   // per-chunk product; chunks are summed in registers
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + kColD2),
                "r"(a), "l"(bd), "r"(idesc2), "r"(acc)
                : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sh.mbar))
                     : "memory");
        mbar_wait(smem_u32(&sh.mbar), phase);
        asm volatile("tcgen05.fence::before_thread_sync;");
      }
      phase ^= 1;
      __syncthreads();   // the chunk's planes and A2 may be overwritten now
      asm volatile("tcgen05.fence::after_thread_sync;");
      {
        // the tensor core's accumulator truncates: keep its sums short (K = 128 per chunk) and add the chunks in FP32
        uint32_t gq[8];
        tmem_ld<8>(lane_addr + kColD2 + 8 * slice, gq);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) gacc[j] += __uint_as_float(gq[j]);
      }
    }
    // ---- g = D2 - theta ; lp = prior + sum over slices of (ll + prior share)
    {
      if (live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int d = 8 * slice + j;
          if (d < D) out_g[(size_t)c * D + d] = gacc[j] - th[j];
        }
      }
      sh.red[slice][row] = ll + prior;
      __syncthreads();
      if (live && slice == 0) out_lp[c] = ((sh.red[0][row] + sh.red[1][row]) + sh.red[2][row]) + sh.red[3][row];
      __syncthreads();
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
  const size_t smem = kChunkBytes;
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
