// Dense-Gaussian log-prob + gradient for ALL chains at once on the 5th-gen tensor cores
// (tcgen05.mma, accumulator in TMEM), FP32-accurate through the 3xTF32 split:
//     G = -(X - mu) P ,  lp = 1/2 <x - mu, g> + c          (MVNTriL.log_prob; gym ill_conditioned_gaussian.py:77-81)
//     X P  ~=  Xhi Phi + Xlo Phi + Xhi Plo ,  hi = tf32(x), lo = tf32(x - hi)
// One CTA = one tile of 128 chains: thread t <-> chain row t <-> TMEM lane t.
//   A operand  : (X - mu) hi/lo planes written by the owning threads straight into TMEM (tcgen05.st)
//   B operand  : P hi/lo planes in shared memory, canonical K-major no-swizzle core-matrix layout
//   D          : [128 x 112] fp32 accumulator in TMEM, A negated by the instruction descriptor => D = g
//   39 MMAs (3 passes x 13 K-steps of 8) of shape M128 N112 K8 per tile, issued by one elected thread,
//   completion signalled through tcgen05.commit -> mbarrier.
#include "pb2_internal.h"

namespace pb2 {

constexpr int kTcM = 128;    // chains per tile
constexpr int kTcNP = 112;   // padded N (D <= 104; N % 16 == 0 for M = 128)
constexpr int kTcKP = 104;   // padded K (multiple of 8)
constexpr int kTcColAhi = 0, kTcColAlo = kTcKP, kTcColD = 2 * kTcKP;   // TMEM columns
constexpr int kTcXs = 101;   // staging row stride (odd: conflict-free row reads)
constexpr int kTcPlaneBytes = (kTcKP / 4) * (kTcNP / 8) * 128;         // 46,592 B per B plane

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// byte offset of element (n, k) of a [kTcNP x kTcKP] K-major operand: 8x(16 B) core matrices,
// K-chunk major (LBO = (NP/8)*128 B between the two K core matrices of one MMA, SBO = 128 B between row groups)
__device__ __forceinline__ int b_plane_offset(int n, int k) {
  return ((k >> 2) * (kTcNP / 8) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 3) * 4;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

__global__ void __launch_bounds__(kTcM, 1)
dense_grad_tc_kernel(const float* __restrict__ X, int B, int D, const float* __restrict__ P,
                     const float* __restrict__ loc, float lognorm, float* __restrict__ lp_out,
                     float* __restrict__ g_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* b_hi = smem_raw;                       // [kTcPlaneBytes]
  unsigned char* b_lo = smem_raw + kTcPlaneBytes;
  float* stage = reinterpret_cast<float*>(smem_raw + 2 * kTcPlaneBytes);   // [128][kTcXs]
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- one-time setup: TMEM allocation, mbarrier, P planes in canonical layout
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = tid; i < kTcNP * kTcKP; i += kTcM) {
    const int n = i / kTcKP, k = i - n * kTcKP;
    const float v = (n < D && k < D) ? P[n * D + k] : 0.f;   // B[n][k] = P[k][n] = P[n][k] (symmetric)
    const float hi = tf32_rna(v);
    const float lo = tf32_rna(v - hi);
    const int off = b_plane_offset(n, k);
    *reinterpret_cast<float*>(b_hi + off) = hi;
    *reinterpret_cast<float*>(b_lo + off) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32-lane quarter
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, A negated, K-major, N=112, M=128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((uint32_t)(kTcNP >> 3) << 17) |
                         ((uint32_t)(kTcM >> 4) << 24);
  const uint32_t lbo = (kTcNP / 8) * 128, sbo = 128;
  const uint32_t bhi_addr = smem_u32(b_hi), blo_addr = smem_u32(b_lo);
  float locv = 0.f;
  uint32_t phase = 0;
  const int ntiles = (B + kTcM - 1) / kTcM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * kTcM;
    const int nrows = min(kTcM, B - row0);
    // ---- coalesced load of the X tile into the staging buffer (row stride 101)
    for (int i = tid; i < nrows * D; i += kTcM) {
      const int r = i / D, c = i - r * D;
      stage[r * kTcXs + c] = X[(size_t)row0 * D + i];
    }
    __syncthreads();
    // ---- split my chain's row into tf32 hi/lo and write both A planes into TMEM (lane = chain)
    const bool live = tid < nrows;
#pragma unroll 1
    for (int c0 = 0; c0 < kTcKP; c0 += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        locv = c < D ? loc[c] : 0.f;
        const float v = (live && c < D) ? (stage[tid * kTcXs + c] - locv) : 0.f;
        const float h = tf32_rna(v);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(tf32_rna(v - h));
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                       lane_addr + kTcColAhi + c0),
                   "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7])
                   : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                       lane_addr + kTcColAlo + c0),
                   "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    // ---- one elected thread issues the 39 MMAs and commits to the mbarrier
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      uint32_t accum = 0;
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a_col = (pass == 1) ? kTcColAlo : kTcColAhi;   // Ahi*Bhi, Alo*Bhi, Ahi*Blo
        const uint32_t b_addr = (pass == 2) ? blo_addr : bhi_addr;
#pragma unroll 1
        for (int j = 0; j < kTcKP / 8; ++j) {
          const uint64_t bdesc = make_kmajor_desc(b_addr + (uint32_t)(2 * j) * lbo, lbo, sbo);
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
              "}\n" ::"r"(tmem + kTcColD),
              "r"(tmem + a_col + 8 * j), "l"(bdesc), "r"(idesc), "r"(accum)
              : "memory");
          accum = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&mbar))
                   : "memory");
    }
    mbar_wait(smem_u32(&mbar), phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    // ---- epilogue: D (= g) back to registers 8 columns at a time; lp = 1/2 <x - mu, g> + c
    float acc = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < kTcKP; c0 += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(lane_addr + kTcColD + c0)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        if (c < D) {
          const float gv = __uint_as_float(v[j]);
          const float xc = stage[tid * kTcXs + c] - loc[c];
          acc = fmaf(xc, gv, acc);
          stage[tid * kTcXs + c] = gv;   // thread t owns row t of the staging buffer
        }
      }
    }
    if (live) lp_out[row0 + tid] = fmaf(0.5f, acc, lognorm);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    for (int i = tid; i < nrows * D; i += kTcM) {
      const int r = i / D, c = i - r * D;
      g_out[(size_t)row0 * D + i] = stage[r * kTcXs + c];
    }
    __syncthreads();
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

int launch_dense_grad_tc(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_x, float* d_lp, float* d_g) {
  const int D = tgt->dim;
  if (D > 100) return set_error(ctx, PB2_ERR_UNSUPPORTED, "tensor-core dense gradient supports D <= 100");
  const size_t smem = 2 * (size_t)kTcPlaneBytes + (size_t)kTcM * kTcXs * sizeof(float);
  if (int rc = check_cuda(ctx, cudaFuncSetAttribute(dense_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)smem), "cudaFuncSetAttribute(dense_grad_tc)"))
    return rc;
  const int ntiles = (B + kTcM - 1) / kTcM;
  const int grid = std::min(ntiles, ctx->num_sms);
  dense_grad_tc_kernel<<<grid, kTcM, smem, ctx->stream>>>(d_x, B, D, tgt->d_a, tgt->d_b, tgt->scalar, d_lp, d_g);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "dense_grad_tc_kernel");
}

}  // namespace pb2

extern "C" int pb2_dense_logp_grad_tc(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_x, float* d_logp,
                                      float* d_grad) {
  using namespace pb2;
  if (!ctx || !tgt || !d_x || !d_logp || !d_grad || B < 1)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_dense_logp_grad_tc: bad argument");
  if (tgt->kind != PB2_TARGET_DENSE_GAUSSIAN)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_dense_logp_grad_tc: target must be a dense Gaussian");
  cudaSetDevice(ctx->device);
  return launch_dense_grad_tc(ctx, tgt, B, d_x, d_logp, d_grad);
}
