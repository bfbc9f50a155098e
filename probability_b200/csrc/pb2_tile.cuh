// Building blocks of the 128-chain TILE kernels for the dense-Gaussian target: the gradient of all
// chains of a tile is ONE tcgen05 contraction  G = -(X - mu) P  (3xTF32 split, FP32 accurate),
//   A = (X - mu) hi/lo planes in TMEM (written by the owning threads with tcgen05.st),
//   B = P hi/lo planes in shared memory (canonical K-major no-swizzle layout),
//   D = [128 x 112] fp32 accumulator in TMEM, negated through the instruction descriptor.
// Thread mapping (512 threads): chain = 32*(warp & 3) + lane  (= its TMEM lane),
// slice = warp >> 2 owns dims [26*slice, 26*slice + 26); per-chain scalars are replicated in
// the 4 threads of a chain and stay bit-identical because cross-slice sums go through
// shared memory in a fixed order.
#pragma once
#include "pb2_internal.h"

namespace pb2 {
namespace tile {

constexpr int kM = 128;       // chains per tile
constexpr int kSlices = 4;
constexpr int kK = 26;        // dims per slice
constexpr int kKP = 104;      // padded K (multiple of 8)
constexpr int kNP = 112;      // padded N (multiple of 16)
constexpr int kThreads = 512;
constexpr int kColAhi = 0, kColAlo = kKP, kColD = 2 * kKP, kColRho = 2 * kKP + kNP;   // TMEM columns (<= 512)
constexpr int kPlaneBytes = (kKP / 4) * (kNP / 8) * 128;
constexpr int kRedN = 8;      // max simultaneous cross-slice reductions

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
  return d;
}

__device__ __forceinline__ int b_plane_offset(int n, int k) {
  return ((k >> 2) * (kNP / 8) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 3) * 4;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TW_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TW_DONE;\n"
      "bra TW_LOOP;\n"
      "TW_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

// ---- TMEM <-> registers, 26 consecutive columns of this thread's lane (x16 + x8 + x2)
__device__ __forceinline__ void tmem_st26(uint32_t a, const uint32_t (&v)[kK]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a + 16), "r"(v[16]),
               "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23])
               : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(a + 24), "r"(v[24]), "r"(v[25]) : "memory");
}

__device__ __forceinline__ void tmem_ld26(uint32_t a, uint32_t (&v)[kK]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(a)
      : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23])
               : "r"(a + 16)
               : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[24]), "=r"(v[25]) : "r"(a + 24) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Shared {
  unsigned long long mbar;
  uint32_t tmem_base;
  uint32_t pad;
  float loc[kKP];
  float red[2][kRedN][kSlices][kM];
};

// Per-thread view of the tile.
struct Ctx {
  Shared* sh;
  uint32_t tmem, lane_addr, bhi_addr, blo_addr, idesc, phase;
  int cl, slice, parity;

  // One-time setup: TMEM allocation, mbarrier, P hi/lo planes (canonical layout), loc.
  __device__ void init(Shared* sh_, unsigned char* planes, const float* P, const float* loc, int D) {
    sh = sh_;
    const int tid = threadIdx.x, warp = tid >> 5;
    cl = 32 * (warp & 3) + (tid & 31);
    slice = warp >> 2;
    parity = 0;
    phase = 0;
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                   "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh->mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;");
    }
    unsigned char* b_hi = planes;
    unsigned char* b_lo = planes + kPlaneBytes;
    for (int i = tid; i < kNP * kKP; i += kThreads) {
      const int n = i / kKP, k = i - n * kKP;
      const float v = (n < D && k < D) ? P[n * D + k] : 0.f;   // B[n][k] = P[k][n] = P[n][k]
      const float hi = tf32_rna(v);
      const int off = b_plane_offset(n, k);
      *reinterpret_cast<float*>(b_hi + off) = hi;
      *reinterpret_cast<float*>(b_lo + off) = tf32_rna(v - hi);
    }
    for (int i = tid; i < kKP; i += kThreads) sh->loc[i] = i < D ? loc[i] : 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    tmem = sh->tmem_base;
    lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    bhi_addr = smem_u32(b_hi);
    blo_addr = smem_u32(b_lo);
    // cute::UMMA::InstrDescriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), negate A (1<<13), K-major, N>>3, M>>4
    idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((uint32_t)(kNP >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
  }

  __device__ void finish() {
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }

  // xc = x - loc of my slice -> tf32 hi/lo planes of the A operand in TMEM
  __device__ __forceinline__ void stage_a(const float (&x)[kK]) {
    uint32_t t[kK];
#pragma unroll
    for (int j = 0; j < kK; ++j) t[j] = __float_as_uint(tf32_rna(x[j] - sh->loc[kK * slice + j]));
    tmem_st26(lane_addr + kColAhi + kK * slice, t);
#pragma unroll
    for (int j = 0; j < kK; ++j) {   // recomputed instead of kept live: registers are the scarce resource
      const float v = x[j] - sh->loc[kK * slice + j];
      t[j] = __float_as_uint(tf32_rna(v - tf32_rna(v)));
    }
    tmem_st26(lane_addr + kColAlo + kK * slice, t);
  }

  // all threads: A is staged -> one thread issues 3 x 13 MMAs (Ahi Bhi + Alo Bhi + Ahi Blo) -> wait
  __device__ __forceinline__ void contract() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t lbo = (kNP / 8) * 128, sbo = 128;
      uint32_t accum = 0;
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a_col = (pass == 1) ? kColAlo : kColAhi;
        const uint32_t b_addr = (pass == 2) ? blo_addr : bhi_addr;
#pragma unroll 1
        for (int j = 0; j < kKP / 8; ++j) {
          const uint64_t bdesc = make_kmajor_desc(b_addr + (uint32_t)(2 * j) * lbo, lbo, sbo);
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
              "}\n" ::"r"(tmem + kColD),
              "r"(tmem + a_col + 8 * j), "l"(bdesc), "r"(idesc), "r"(accum)
              : "memory");
          accum = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sh->mbar))
                   : "memory");
    }
    mbar_wait(smem_u32(&sh->mbar), phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
  }

  // my slice of D (= gradient)
  __device__ __forceinline__ void load_d(float (&g)[kK]) {
    uint32_t t[kK];
    tmem_ld26(lane_addr + kColD + kK * slice, t);
#pragma unroll
    for (int j = 0; j < kK; ++j) g[j] = __uint_as_float(t[j]);
  }

  // cross-slice sums (fixed order => the 4 threads of a chain get identical bits); one barrier
  template <int N>
  __device__ __forceinline__ void reduce(float (&v)[N]) {
    static_assert(N <= kRedN, "too many simultaneous reductions");
    float(*buf)[kSlices][kM] = sh->red[parity];
    parity ^= 1;
#pragma unroll
    for (int i = 0; i < N; ++i) buf[i][slice][cl] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = ((buf[i][0][cl] + buf[i][1][cl]) + buf[i][2][cl]) + buf[i][3][cl];
  }
};

}  // namespace tile
}  // namespace pb2
