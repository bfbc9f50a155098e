// Building blocks of the 128-chain TILE kernels (tile_hmc_kernel, pb2_tile.cu; the NUTS kernels use the 64-chain
// variant pb2_tile64.cuh, which reuses the helpers below) for the dense-Gaussian target: the gradient of all
// chains of a tile is ONE tcgen05 contraction  G = -(X - mu) P  (3xTF32 split, FP32 accurate),
//   A = (X - mu) hi/lo planes in TMEM (written by the owning threads with tcgen05.st),
//   B = P hi/lo planes in shared memory (canonical K-major no-swizzle layout),
//   D = [128 x 112] fp32 accumulator in TMEM, negated through the instruction descriptor.
// Thread mapping (512 threads): chain = 32*(warp & 3) + lane  (= its TMEM lane),
// slice = warp >> 2 owns dims [26*slice, 26*slice + 26); per-chain scalars are replicated in
// the 4 threads of a chain and stay bit-identical because cross-slice sums go through
// shared memory in a fixed order.
#pragma once
#include <type_traits>
#include "pb2_internal.h"

namespace pb2 {
namespace tile {

constexpr int kM = 128;       // chains per tile
constexpr int kSlices = 4;
constexpr int kK = 26;        // dims per slice
constexpr int kKP = 104;      // padded K (multiple of 8)
constexpr int kNP = 112;      // padded N (multiple of 16)
constexpr int kThreads = 512;
constexpr int kColAhi = 0, kColAlo = kKP, kColD = 2 * kKP;   // TMEM columns (<= 512)
constexpr int kPlaneBytes = (kKP / 4) * (kNP / 8) * 128;
constexpr int kRedN = 4;      // max simultaneous cross-slice reductions

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// cute::UMMA::SmemDescriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version(1) <<46, no swizzle
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// byte offset of element (n, k) of a [kNP x kKP] K-major operand made of 8 x 16 B core matrices,
// K-chunk major: LBO = (NP/8)*128 B between the two K core matrices of one MMA, SBO = 128 B between row groups
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

// ---- TMEM <-> registers: N consecutive 32-bit columns of this thread's lane
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t a, const uint32_t (&v)[N]) {
  static_assert(N == 16 || N == 8 || N == 4 || N == 2 || N == 1, "chunk sizes used by the tile kernels");
  if constexpr (N == 4) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
                 : "memory");
  } else if constexpr (N == 1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a), "r"(v[0]) : "memory");
  } else if constexpr (N == 16) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
  } else if constexpr (N == 8) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
  } else {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(a), "r"(v[0]), "r"(v[1]) : "memory");
  }
}

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t a, uint32_t (&v)[N]) {   // caller issues tcgen05.wait::ld
  static_assert(N == 16 || N == 8 || N == 4 || N == 2 || N == 1, "chunk sizes used by the tile kernels");
  if constexpr (N == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(a)
                 : "memory");
  } else if constexpr (N == 1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(a) : "memory");
  } else if constexpr (N == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(a)
        : "memory");
  } else if constexpr (N == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(a)
                 : "memory");
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(a) : "memory");
  }
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// compile-time loop over the chunks (offset, length) that tile the 26 columns of a slice
template <class F>
__device__ __forceinline__ void for_chunks(F&& f) {
  f(std::integral_constant<int, 0>{}, std::integral_constant<int, 16>{});
  f(std::integral_constant<int, 16>{}, std::integral_constant<int, 8>{});
  f(std::integral_constant<int, 24>{}, std::integral_constant<int, 2>{});
}

struct Shared {
  unsigned long long mbar;
  uint32_t tmem_base;
  int flags[4];   // "some chain still continues" flags, rotated per leaf / per doubling
  float loc[kKP];
  float red[2][kRedN][kSlices][kM];
};

// Per-thread view of the tile.
struct Ctx {
  Shared* sh;
  uint32_t tmem, lane_addr, idesc, phase;
  uint64_t bdesc_hi, bdesc_lo;
  int cl, slice, parity;

  // One-time setup: TMEM allocation, mbarrier, P hi/lo planes (canonical layout), loc.
  __device__ void init(Shared* sh_, unsigned char* planes, const float* P, const float* loc, int D, const float* scale = nullptr) {
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
      sh->flags[0] = sh->flags[1] = sh->flags[2] = sh->flags[3] = 0;
    }
    unsigned char* b_hi = planes;
    unsigned char* b_lo = planes + kPlaneBytes;
    for (int i = tid; i < kNP * kKP; i += kThreads) {
      const int n = i / kKP, k = i - n * kKP;
      // B[n][k] = P[k][n] = P[n][k]; a diagonally preconditioned run samples u = x / s: precision diag(s) P diag(s)
      float v = (n < D && k < D) ? P[n * D + k] : 0.f;
      if (scale && n < D && k < D) v = (scale[n] * v) * scale[k];
      const float hi = tf32_rna(v);
      const int off = b_plane_offset(n, k);
      *reinterpret_cast<float*>(b_hi + off) = hi;
      *reinterpret_cast<float*>(b_lo + off) = tf32_rna(v - hi);
    }
    for (int i = tid; i < kKP; i += kThreads) sh->loc[i] = i < D ? (scale ? loc[i] / scale[i] : loc[i]) : 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    tmem = sh->tmem_base;
    lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    bdesc_hi = make_kmajor_desc(smem_u32(b_hi), (kNP / 8) * 128, 128);
    bdesc_lo = make_kmajor_desc(smem_u32(b_lo), (kNP / 8) * 128, 128);
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
    const uint32_t base = lane_addr + kK * slice;
    const float* lc = sh->loc + kK * slice;
    for_chunks([&](auto off, auto n) {
      constexpr int OFF = decltype(off)::value, N = decltype(n)::value;
      uint32_t hi[N], lo[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        // round-to-nearest tf32 split with integer ops (cvt.rna.tf32.f32 is emulated with ~5 instructions on
        // sm_100a: inf/nan handling this path does not need -- a non-finite x gives a NaN energy = divergence)
        const float v = x[OFF + j] - lc[OFF + j];
        hi[j] = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
        lo[j] = (__float_as_uint(v - __uint_as_float(hi[j])) + 0x1000u) & 0xffffe000u;
      }
      tmem_st<N>(base + kColAhi + OFF, hi);
      tmem_st<N>(base + kColAlo + OFF, lo);
    });
  }

  // The issue loop runs on ONE thread, so every instruction in it is on the critical path of the
  // tick: fully unrolled, descriptors precomputed.
  template <int PASS, int J>
  __device__ __forceinline__ void issue_one() {
    constexpr uint32_t a_col = (PASS == 1) ? kColAlo : kColAhi;           // Ahi*Bhi, Alo*Bhi, Ahi*Blo
    constexpr uint32_t step = (2u * (kNP / 8) * 128u) >> 4;               // two K core matrices per MMA (16 B units)
    const uint64_t bdesc = ((PASS == 2) ? bdesc_lo : bdesc_hi) + (uint64_t)(J * step);
    if constexpr (PASS == 0 && J == 0) {
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, 0, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + kColD),
          "r"(tmem + a_col + 8 * J), "l"(bdesc), "r"(idesc)
          : "memory");
    } else {
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + kColD),
          "r"(tmem + a_col + 8 * J), "l"(bdesc), "r"(idesc)
          : "memory");
    }
  }
  template <int PASS, int J>
  __device__ __forceinline__ void issue_from() {
    issue_one<PASS, J>();
    if constexpr (J + 1 < kKP / 8) issue_from<PASS, J + 1>();
    else if constexpr (PASS + 1 < 3) issue_from<PASS + 1, 0>();
  }

  // all threads: A is staged -> thread 0 issues 3 x 13 MMAs (M128 N112 K8) + commit -> everybody waits
  __device__ __forceinline__ void contract() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      issue_from<0, 0>();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sh->mbar))
                   : "memory");
      mbar_wait(smem_u32(&sh->mbar), phase);   // only the issuing thread polls; everybody else sleeps in bar.sync
      asm volatile("tcgen05.fence::before_thread_sync;");
    }
    phase ^= 1;
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }

  // my slice of D (= gradient)
  __device__ __forceinline__ void load_d(float (&g)[kK]) {
    const uint32_t base = lane_addr + kColD + kK * slice;
    uint32_t t0[16], t1[8], t2[2];
    tmem_ld<16>(base, t0);
    tmem_ld<8>(base + 16, t1);
    tmem_ld<2>(base + 24, t2);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) g[j] = __uint_as_float(t0[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[16 + j] = __uint_as_float(t1[j]);
    g[24] = __uint_as_float(t2[0]);
    g[25] = __uint_as_float(t2[1]);
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
