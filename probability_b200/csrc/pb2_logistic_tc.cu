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
// The chunk operands (four pre-split planes, one contiguous block per chunk) are streamed by the TMA: one
// cp.async.bulk.tensor per chunk (the plane buffer described as a 2-d tensor of 1 KB rows, box = one chunk) into a
// 4-slot ring, completion on the slot's mbarrier, which the issuing warp waits on before the chunk's first contraction.
// This is the primitive (SURVEY 8a row T3 on tcgen05); the transition kernels still use the FP32 warp-per-chain
// gradient (pb2_targets.cuh LogisticT).
#include <cuda.h>   // CUtensorMap (types only: cuTensorMapEncodeTiled is resolved through the runtime at launch)

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
constexpr int kWorkers = 512;  // 16 worker warps: thread = (chain row, slice = warp >> 2)
constexpr int kThreads = kWorkers + 32;   // + warp 16: issues the contractions
constexpr int kRing = 4;       // chunk operand buffers: GEMM 2 of c - 1 and c, GEMM 1 of c + 1, async copy of c + 2

// Shapes.  KD = padded D as the K of GEMM 1 (multiple of 8), ND = padded D as the N of GEMM 2 (multiple of 16),
// R = rows of X~ per chunk (N of GEMM 1, K of GEMM 2).  TMEM columns: theta hi/lo | 2 x z chunk | 2 x r hi |
// 2 x r lo | g chunk  =  2 KD + 6 R + ND <= 512.
template <int KD_, int ND_, int R_>
struct Shape {
  static constexpr int KD = KD_, ND = ND_, R = R_;
  static constexpr int kColA1hi = 0, kColA1lo = KD, kColD1 = 2 * KD, kColA2hi = 2 * KD + 2 * R,
                       kColA2lo = 2 * KD + 4 * R, kColD2 = 2 * KD + 6 * R;
  static_assert(kColD2 + ND <= 512, "TMEM budget");
  static constexpr int kB1Plane = R * KD * 4;   // [n = R rows][k = KD dims]
  static constexpr int kB2Plane = ND * R * 4;   // [n = ND dims][k = R rows]
  static constexpr int kChunkBytes = 2 * kB1Plane + 2 * kB2Plane;   // hi/lo of both layouts
  static constexpr int kTmaRowBytes = 1024;                         // the TMA moves a chunk as rows of 1 KB
  static constexpr int kChunkRows = kChunkBytes / kTmaRowBytes;
  static_assert(kChunkBytes % kTmaRowBytes == 0 && kChunkRows <= 256, "TMA box");
  static constexpr int kTh = KD / 4;            // theta dims per worker thread
  static constexpr int kG = ND / 4;             // gradient columns per worker thread
  static constexpr int kZ = R / 4;              // logits per worker thread and chunk
  // the tensor core's accumulator truncates: it only sums kGroup chunks (128 rows) before the workers take the
  // partial gradient out of D2 and add it in FP32 registers
  static constexpr int kGroup = 128 / R;
};
using SmallD = Shape<32, 32, 64>;      // D <= 32 (C3: 1000 x 25): 480 columns, 32 KB per chunk
using LargeD = Shape<104, 112, 32>;    // D <= 100 (C5: 1e6 x 100): 512 columns, 54 KB per chunk

// byte offset of element (n, k) of a [NP x KP] K-major no-swizzle operand made of 8 x 16 B core matrices, K-chunk
// major: LBO = (NP/8)*128 B between the two K core matrices of one MMA, SBO = 128 B between row groups
__host__ __device__ inline int plane_offset(int NP, int n, int k) {
  return ((k >> 2) * (NP / 8) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 3) * 4;
}

__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ uint32_t tf32_round(float v) { return (__float_as_uint(v) + 0x1000u) & 0xffffe000u; }

// N consecutive TMEM columns of my lane <-> registers, N split greedily into the x16 / x8 / x4 / x2 / x1 shapes
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t a, uint32_t* v) {
  if constexpr (N >= 16) { uint32_t t[16]; tmem_ld<16>(a, t);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = t[j];
    if constexpr (N > 16) tmem_ld_n<N - 16>(a + 16, v + 16);
  } else if constexpr (N >= 8) { uint32_t t[8]; tmem_ld<8>(a, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = t[j];
    if constexpr (N > 8) tmem_ld_n<N - 8>(a + 8, v + 8);
  } else if constexpr (N >= 4) { uint32_t t[4]; tmem_ld<4>(a, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = t[j];
    if constexpr (N > 4) tmem_ld_n<N - 4>(a + 4, v + 4);
  } else if constexpr (N >= 2) { uint32_t t[2]; tmem_ld<2>(a, t); v[0] = t[0]; v[1] = t[1];
    if constexpr (N > 2) tmem_ld_n<N - 2>(a + 2, v + 2);
  } else { uint32_t t[1]; tmem_ld<1>(a, t); v[0] = t[0]; }
}
template <int N>
__device__ __forceinline__ void tmem_st_n(uint32_t a, const uint32_t* v) {
  if constexpr (N >= 16) { uint32_t t[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = v[j];
    tmem_st<16>(a, t);
    if constexpr (N > 16) tmem_st_n<N - 16>(a + 16, v + 16);
  } else if constexpr (N >= 8) { uint32_t t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[j];
    tmem_st<8>(a, t);
    if constexpr (N > 8) tmem_st_n<N - 8>(a + 8, v + 8);
  } else if constexpr (N >= 4) { uint32_t t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = v[j];
    tmem_st<4>(a, t);
    if constexpr (N > 4) tmem_st_n<N - 4>(a + 4, v + 4);
  } else if constexpr (N >= 2) { uint32_t t[2] = {v[0], v[1]}; tmem_st<2>(a, t);
    if constexpr (N > 2) tmem_st_n<N - 2>(a + 2, v + 2);
  } else { uint32_t t[1] = {v[0]}; tmem_st<1>(a, t); }
}

// ---- one-time: X~ [N, ldx] -> per chunk the four planes (B1 hi, B1 lo, B2 hi, B2 lo) in the canonical layouts
template <class S>
__global__ void logistic_tc_prepare_kernel(const float* __restrict__ X, int N, int D, int ldx,
                                           unsigned char* __restrict__ out) {
  const int chunk = blockIdx.x;
  unsigned char* o = out + (size_t)chunk * S::kChunkBytes;
  constexpr int DD = S::KD > S::ND ? S::KD : S::ND;
  for (int i = threadIdx.x; i < S::R * DD; i += blockDim.x) {
    const int r = i / DD, d = i - r * DD;   // row of the chunk, dim
    const int n = chunk * S::R + r;
    const float v = (n < N && d < D) ? X[(size_t)n * ldx + d] : 0.f;
    const uint32_t hi = tf32_round(v);
    const uint32_t lo = tf32_round(v - __uint_as_float(hi));
    if (d < S::KD) {
      const int o1 = plane_offset(S::R, r, d);     // B1[n = r][k = d]
      *reinterpret_cast<uint32_t*>(o + o1) = hi;
      *reinterpret_cast<uint32_t*>(o + S::kB1Plane + o1) = lo;
    }
    if (d < S::ND) {
      const int o2 = plane_offset(S::ND, d, r);    // B2[n = d][k = r]
      *reinterpret_cast<uint32_t*>(o + 2 * S::kB1Plane + o2) = hi;
      *reinterpret_cast<uint32_t*>(o + 2 * S::kB1Plane + S::kB2Plane + o2) = lo;
    }
  }
}

template <class S>
struct Smem {
  unsigned long long g1_done[2], g2_done;   // mbarriers: the contraction into D1[b] / D2 has completed
  unsigned long long full[kRing];           // mbarriers: the TMA has delivered the ring slot's operand planes
  uint32_t tmem_base;
  float y[kRing][S::R];
  float valid[kRing][S::R];
  float red[4][kM];
};

// Software pipeline (per 128-chain tile, chunk c of R rows of X~):
//   workers : start the async copy of chunk c+2's operands -> wait z(c) -> read it (D1[c%2] is free) -> sigmoid /
//             softplus of my logits -> wait for GEMM 2 of c-1 (take the g group out of D2) -> signal A -> r hi/lo into
//             A2[c%2] -> signal B
//   issuer  : on A: GEMM 1 of chunk c+2 into D1[c%2];  on B: GEMM 2 of chunk c into D2
// so GEMM 1 of c+2 and GEMM 2 of c run back to back in the shadow of the workers' transcendental work on chunk c+1.  UTCHMMA issue is back-pressured by
// the tensor pipe, which is why it lives in a warp of its own.  Signals A / B are named barriers 2 / 3 on which the
// workers only arrive.
// kPartial = false: one CTA runs whole tiles over ALL rows and writes log-prob and gradient (prior included).
// kPartial = true : CTA (tile = blockIdx.x, segment = blockIdx.y) covers `seg_chunks` chunks of the rows and writes its
//                   partial sums part_g[segment][B][D], part_ll[segment][B] (row-sharded data: the caller reduces).
template <class S, bool kPartial>
__global__ void __launch_bounds__(kThreads, 1)
logistic_tc_kernel(const float* __restrict__ Theta, int B, int D, int N, const __grid_constant__ CUtensorMap planes_map,
                   const float* __restrict__ labels, int nchunks_total, int seg_chunks, float* __restrict__ out_lp,
                   float* __restrict__ out_g) {
  extern __shared__ __align__(128) unsigned char ring[];   // kRing x (B1 hi, B1 lo, B2 hi, B2 lo)
  __shared__ Smem<S> sh;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.g1_done[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.g1_done[1])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.g2_done)));
    for (int q = 0; q < kRing; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh.full[q])));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&planes_map) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = sh.tmem_base;
  const int ntiles = (B + kM - 1) / kM;
  // my tiles and my chunk range
  const int tile0 = blockIdx.x, tile_step = kPartial ? ntiles : gridDim.x;   // kPartial: exactly one tile per CTA
  const int cbeg = kPartial ? blockIdx.y * seg_chunks : 0;
  const int cend = kPartial ? min(nchunks_total, cbeg + seg_chunks) : nchunks_total;
  const int nchunks = max(0, cend - cbeg);
  const int my_tiles = tile0 < ntiles ? (kPartial ? 1 : (ntiles - 1 - tile0) / tile_step + 1) : 0;

  if (warp == kWorkers / 32) {
    // ------------------------------------------------------------------ the issuing warp
    const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(S::R >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(S::ND >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    uint32_t leader;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
    uint32_t full_parity = 0u;   // bit q: phase parity of ring slot q's `full` barrier (a register, not an indexed array)
    auto gemm1 = [&](int c) {   // z chunk = theta . X~chunk^T : 3 passes x KD/8 K-steps, M128 N=R K8
      if (!leader) return;
      mbar_wait(smem_u32(&sh.full[c % kRing]), (full_parity >> (c % kRing)) & 1u);   // the chunk's planes have landed
      full_parity ^= 1u << (c % kRing);
      const uint32_t base = smem_u32(ring + (size_t)(c % kRing) * S::kChunkBytes);
      const uint64_t bhi = make_kmajor_desc(base, (S::R / 8) * 128, 128);
      const uint64_t blo = make_kmajor_desc(base + S::kB1Plane, (S::R / 8) * 128, 128);
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
        for (int j = 0; j < S::KD / 8; ++j) {
          const uint32_t a = tmem + (pass == 1 ? S::kColA1lo : S::kColA1hi) + 8 * j;
          const uint64_t bd = (pass == 2 ? blo : bhi) + (uint64_t)(j * ((2u * (S::R / 8) * 128u) >> 4));
          const uint32_t acc = (pass == 0 && j == 0) ? 0u : 1u;
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + S::kColD1 + S::R * (c & 1)),
              "r"(a), "l"(bd), "r"(idesc1), "r"(acc)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&sh.g1_done[c & 1]))
                   : "memory");
    };
    auto gemm2 = [&](int c) {   // g chunk = r . X~chunk : 3 passes x R/8 K-steps, M128 N=ND K8
      if (!leader) return;
      const uint32_t base = smem_u32(ring + (size_t)(c % kRing) * S::kChunkBytes) + 2 * S::kB1Plane;
      const uint64_t bhi = make_kmajor_desc(base, (S::ND / 8) * 128, 128);
      const uint64_t blo = make_kmajor_desc(base + S::kB2Plane, (S::ND / 8) * 128, 128);
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
        for (int j = 0; j < S::R / 8; ++j) {
          const uint32_t a = tmem + (pass == 1 ? S::kColA2lo : S::kColA2hi) + S::R * (c & 1) + 8 * j;
          const uint64_t bd = (pass == 2 ? blo : bhi) + (uint64_t)(j * ((2u * (S::ND / 8) * 128u) >> 4));
          const uint32_t acc = (c % S::kGroup == 0 && pass == 0 && j == 0) ? 0u : 1u;   // D2 sums kGroup chunks
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + S::kColD2),
              "r"(a), "l"(bd), "r"(idesc2), "r"(acc)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sh.g2_done))
                   : "memory");
    };
    for (int tl = 0; tl < my_tiles; ++tl) {
      if (nchunks == 0) break;
      asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");   // theta staged, chunks 0 and 1 requested
      asm volatile("tcgen05.fence::after_thread_sync;");
      gemm1(0);
      if (nchunks > 1) gemm1(1);
      for (int c = 0; c < nchunks; ++c) {
        asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");   // D1[c%2] read, D2 read, chunk c+2 requested
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
    const int slice = warp >> 2;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t g1_parity = 0u, n2 = 0u;               // bit b: phase parity of g1_done[b]; n2: completed g2_done waits
    auto stage = [&](int c) {   // chunk cbeg + c's operand planes (one TMA tensor copy) and labels -> ring slot c % kRing
      if (tid == 0) {
        const uint32_t dst = smem_u32(ring + (size_t)(c % kRing) * S::kChunkBytes);
        const uint32_t bar = smem_u32(&sh.full[c % kRing]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(S::kChunkBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
            "l"(&planes_map), "r"(0), "r"((cbeg + c) * S::kChunkRows), "r"(bar)
            : "memory");
      }
      if (tid < S::R) {
        const int n = (cbeg + c) * S::R + tid;
        sh.y[c % kRing][tid] = n < N ? labels[n] : 0.f;
        sh.valid[c % kRing][tid] = n < N ? 1.f : 0.f;
      }
    };
    auto signal = [&](int bar) {   // workers only ARRIVE (after making their TMEM writes visible)
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      if (bar == 2) asm volatile("bar.arrive 2, %0;" ::"n"(kThreads) : "memory");
      else asm volatile("bar.arrive 3, %0;" ::"n"(kThreads) : "memory");
    };
    for (int tl = 0; tl < my_tiles && nchunks > 0; ++tl) {
      const int tile_i = tile0 + tl * tile_step;
      const int c0 = tile_i * kM + row;
      const bool live = c0 < B;
      // ---- A1 = theta hi/lo: my KD/4 dims (slice) of my chain
      float th[S::kTh];
      float prior = 0.f;
      {
        uint32_t hi[S::kTh], lo[S::kTh];
#pragma unroll
        for (int j = 0; j < S::kTh; ++j) {
          const int d = S::kTh * slice + j;
          th[j] = (live && d < D) ? Theta[(size_t)c0 * D + d] : 0.f;
          if (d < D) prior += -0.5f * th[j] * th[j] - 0.9189385332046727f;
          hi[j] = tf32_round(th[j]);
          lo[j] = tf32_round(th[j] - __uint_as_float(hi[j]));
        }
        tmem_st_n<S::kTh>(lane_addr + S::kColA1hi + S::kTh * slice, hi);
        tmem_st_n<S::kTh>(lane_addr + S::kColA1lo + S::kTh * slice, lo);
      }
      stage(0);
      if (nchunks > 1) stage(1);
      signal(2);
      float ll = 0.f;   // my share of sum_n [y z - softplus z]
      float gacc[S::kG];
#pragma unroll
      for (int j = 0; j < S::kG; ++j) gacc[j] = 0.f;
      auto take_g = [&](bool read) {   // contraction 2 of the previous chunk: wait; at the end of a group of chunks
        mbar_wait(smem_u32(&sh.g2_done), n2 & 1u);   // take the partial gradient out of D2 and add it in FP32
        n2++;
        if (!read) return;
        asm volatile("tcgen05.fence::after_thread_sync;");
        uint32_t gq[S::kG];
        tmem_ld_n<S::kG>(lane_addr + S::kColD2 + S::kG * slice, gq);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < S::kG; ++j) gacc[j] += __uint_as_float(gq[j]);
      };
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        if (c + 2 < nchunks) stage(c + 2);   // slot (c+2) % 4 was last read by chunk c-2's contractions (waited for)
        mbar_wait(smem_u32(&sh.g1_done[b]), (g1_parity >> b) & 1u);
        g1_parity ^= 1u << b;
        asm volatile("tcgen05.fence::after_thread_sync;");
        uint32_t zq[S::kZ];
        tmem_ld_n<S::kZ>(lane_addr + S::kColD1 + S::R * b + S::kZ * slice, zq);
        tmem_wait_ld();
        // ---- my logits -> log-likelihood terms, r = y - sigmoid(z) -> A2[b] hi/lo (chunk c-1's contraction 2 runs
        // on the tensor pipe meanwhile: it is only waited for below, after the transcendental work)
        uint32_t hi[S::kZ], lo[S::kZ];
        const float* yy = sh.y[c % kRing] + S::kZ * slice;
        const float* vv = sh.valid[c % kRing] + S::kZ * slice;
#pragma unroll
        for (int j = 0; j < S::kZ; ++j) {
          const float z = __uint_as_float(zq[j]);
          // softplus(z) = max(z, 0) + log1p(exp(-|z|)); sigmoid(z) from the same exponential
          // the fast forms of the FP32 kernel (pb2_targets.cuh: __expf / __logf / __fdividef) without their
          // denormal-range fix-ups: e below 2^-126 is flushed, where 1 + e == 1 anyway, and 1 + e is in [1, 2]
          const float e = ex2_ftz(fabsf(z) * -1.4426950216293334961f);
          const float sp = fmaxf(z, 0.f) + lg2_ftz(1.0f + e) * 0.693147182464599609375f;
          const float inv = rcp_ftz(1.0f + e);
          const float sg = z >= 0.f ? inv : e * inv;
          ll += vv[j] * (yy[j] * z - sp);
          const float r = vv[j] * (yy[j] - sg);
          hi[j] = tf32_round(r);
          lo[j] = tf32_round(r - __uint_as_float(hi[j]));
        }
        // chunk c-1's contraction 2 is the last reader of A2[(c-1)%2] and the writer of D2; waiting for it before
        // signal A also keeps every worker within one round of the issuer (the workers only ARRIVE on barriers 2 / 3)
        if (c >= 1) take_g(c % S::kGroup == 0);
        signal(2);
        tmem_st_n<S::kZ>(lane_addr + S::kColA2hi + S::R * b + S::kZ * slice, hi);
        tmem_st_n<S::kZ>(lane_addr + S::kColA2lo + S::R * b + S::kZ * slice, lo);
        signal(3);
      }
      take_g(true);
      if constexpr (kPartial) {
        // ---- partial sums of my row segment
        float* pg = out_g + ((size_t)blockIdx.y * B + c0) * D;
        if (live) {
#pragma unroll
          for (int j = 0; j < S::kG; ++j) {
            const int d = S::kG * slice + j;
            if (d < D) pg[d] = gacc[j];
          }
        }
        sh.red[slice][row] = ll;
        asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
        if (live && slice == 0)
          out_lp[(size_t)blockIdx.y * B + c0] = ((sh.red[0][row] + sh.red[1][row]) + sh.red[2][row]) + sh.red[3][row];
        asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
      } else {
        // ---- g = sum of chunks - theta ; lp = prior + log-likelihood, summed over the 4 slices in a fixed order
        // (gradient columns and theta dims are sliced alike only if kG == kTh: true for SmallD)
        static_assert(kPartial || S::kG == S::kTh, "full mode needs matching theta / gradient slices");
        if (live) {
#pragma unroll
          for (int j = 0; j < S::kG; ++j) {
            const int d = S::kG * slice + j;
            if (d < D) out_g[(size_t)c0 * D + d] = gacc[j] - th[j];
          }
        }
        sh.red[slice][row] = ll + prior;
        asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
        if (live && slice == 0) out_lp[c0] = ((sh.red[0][row] + sh.red[1][row]) + sh.red[2][row]) + sh.red[3][row];
        asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// packed[b][0..D) = sum over segments of part_g, packed[b][D] = sum of part_ll (fixed order: deterministic)
__global__ void logistic_tc_reduce_kernel(const float* __restrict__ part_g, const float* __restrict__ part_ll, int S, int B,
                                          int D, float* __restrict__ packed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * (D + 1)) return;
  const int b = (int)(i / (D + 1)), d = (int)(i - (size_t)b * (D + 1));
  float acc = 0.f;
  for (int s = 0; s < S; ++s) acc += d < D ? part_g[((size_t)s * B + b) * D + d] : part_ll[(size_t)s * B + b];
  packed[i] = acc;
}
}  // namespace ltc

// The plane buffer (nchunks blocks of S::kChunkBytes) as a 2-d float tensor of 1 KB rows; box = one chunk.
template <class S>
static int make_planes_map(pb2_ctx* ctx, const void* d_planes, int nchunks, CUtensorMap* map) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (int rc = check_cuda(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres),
                            "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)"))
      return rc;
    if (!fn || qres != cudaDriverEntryPointSuccess)
      return set_error(ctx, PB2_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)(S::kTmaRowBytes / 4), (cuuint64_t)std::max(nchunks, 1) * S::kChunkRows};
  const cuuint64_t gstride[1] = {(cuuint64_t)S::kTmaRowBytes};
  const cuuint32_t box[2] = {(cuuint32_t)(S::kTmaRowBytes / 4), (cuuint32_t)S::kChunkRows};
  const cuuint32_t estride[2] = {1u, 1u};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(d_planes), gdim, gstride, box, estride,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(ctx, PB2_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return PB2_OK;
}

int launch_logistic_tc(pb2_ctx* ctx, pb2_target* tgt, int B, const float* d_x, float* d_lp, float* d_g) {
  using namespace ltc;
  using S = SmallD;
  const int D = tgt->dim, N = tgt->n_rows;
  if (D > S::KD) return set_error(ctx, PB2_ERR_UNSUPPORTED, "pb2_logistic_logp_grad_tc: D <= 32 only");
  const int nchunks = (N + S::R - 1) / S::R;
  const size_t need = (size_t)nchunks * S::kChunkBytes;
  if (!tgt->d_tc) {   // the operand planes of this target, built once
    if (int rc = check_cuda(ctx, cudaMalloc(&tgt->d_tc, need), "cudaMalloc(logistic tc planes)")) return rc;
    tgt->tc_bytes = need;
    logistic_tc_prepare_kernel<S><<<nchunks, 256, 0, ctx->stream>>>(tgt->d_a, N, D, D, tgt->d_tc);
    ctx->launches += 1;
  }
  const size_t smem = (size_t)kRing * S::kChunkBytes;
  auto kern = logistic_tc_kernel<S, false>;
  if (int rc = check_cuda(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(logistic_tc)"))
    return rc;
  const int grid = std::min((B + kM - 1) / kM, ctx->num_sms);
  CUtensorMap map;
  if (int rc = make_planes_map<S>(ctx, tgt->d_tc, nchunks, &map)) return rc;
  kern<<<grid, kThreads, smem, ctx->stream>>>(d_x, B, D, N, map, tgt->d_b, nchunks, nchunks, d_lp, d_g);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "logistic_tc_kernel");
}

// row-sharded data (C5): this rank's rows as operand planes (built once by launch_rowshard_tc_prepare into a
// caller-owned buffer), all chains; packed [B, D + 1]
size_t rowshard_tc_planes_bytes(int N) {
  using S = ltc::LargeD;
  return (size_t)std::max((N + S::R - 1) / S::R, 1) * S::kChunkBytes;
}

int launch_rowshard_tc_prepare(pb2_ctx* ctx, const float* d_X, int N, int D, int DP, unsigned char* d_planes) {
  using namespace ltc;
  using S = LargeD;
  if (D > 100) return set_error(ctx, PB2_ERR_UNSUPPORTED, "pb2_rowshard_tc_prepare: D <= 100 only");
  const int nchunks = (N + S::R - 1) / S::R;
  if (nchunks > 0) {
    logistic_tc_prepare_kernel<S><<<nchunks, 256, 0, ctx->stream>>>(d_X, N, D, DP, d_planes);
    ctx->launches += 1;
  }
  return check_cuda(ctx, cudaGetLastError(), "logistic_tc_prepare_kernel");
}

int launch_rowshard_tc(pb2_ctx* ctx, const unsigned char* d_planes, const float* d_y, int N, int D, const float* d_theta,
                       int B, float* d_packed) {
  using namespace ltc;
  using S = LargeD;
  if (D > 100) return set_error(ctx, PB2_ERR_UNSUPPORTED, "pb2_rowshard_logistic_grad_tc: D <= 100 only");
  const int nchunks = (N + S::R - 1) / S::R;
  const int ntiles = (B + kM - 1) / kM;
  int nseg = std::max(1, ctx->num_sms / ntiles);
  nseg = std::min(nseg, std::max(1, nchunks));
  const int seg_chunks = (std::max(nchunks, 1) + nseg - 1) / nseg;
  nseg = (std::max(nchunks, 1) + seg_chunks - 1) / seg_chunks;
  const size_t pneed = sizeof(float) * ((size_t)nseg * B * D + (size_t)nseg * B);
  if (pneed > ctx->sched_bytes) {
    if (ctx->d_sched) cudaFree(ctx->d_sched);
    ctx->d_sched = nullptr;
    ctx->sched_bytes = 0;
    if (int rc = check_cuda(ctx, cudaMalloc((void**)&ctx->d_sched, pneed), "cudaMalloc(rowshard partials)")) return rc;
    ctx->sched_bytes = pneed;
  }
  float* part_g = reinterpret_cast<float*>(ctx->d_sched);
  float* part_ll = part_g + (size_t)nseg * B * D;
  // every (tile, segment) CTA overwrites its whole block of the partials (all segments hold >= 1 chunk); only an empty
  // shard leaves them untouched
  if (nchunks == 0)
    if (int rc = check_cuda(ctx, cudaMemsetAsync(part_g, 0, pneed, ctx->stream), "memset(rowshard partials)")) return rc;
  const size_t smem = (size_t)kRing * S::kChunkBytes;
  auto kern = logistic_tc_kernel<S, true>;
  if (int rc = check_cuda(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(rowshard_tc)"))
    return rc;
  if (nchunks > 0) {
    CUtensorMap map;
    if (int rc = make_planes_map<S>(ctx, d_planes, nchunks, &map)) return rc;
    kern<<<dim3(ntiles, nseg), kThreads, smem, ctx->stream>>>(d_theta, B, D, N, map, d_y, nchunks, seg_chunks, part_ll, part_g);
    ctx->launches += 1;
  }
  const size_t tot = (size_t)B * (D + 1);
  logistic_tc_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(part_g, part_ll, nseg, B, D, d_packed);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "rowshard_tc kernels");
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

// SimpleLeapfrogIntegrator (leapfrog_integrator.py:280-355) for the logistic-regression target with ALL chains in
// lock-step: every leapfrog's log-prob + gradient is one launch of the tcgen05 kernel above, between them one fused
// kick + drift kernel -- the tensor-core transition path of T3 (HMC: a fixed L, so lock-step wastes nothing).
extern "C" int pb2_lockstep_leapfrog(pb2_ctx* ctx, int mode, int B, int D, const float* d_step, int step_kind, float* d_v,
                                     float* d_x, const float* d_g, const float* d_m_in, float* d_m_out);
extern "C" int pb2_logistic_tc_leapfrog(pb2_ctx* ctx, const pb2_target* tgt, int B, const float* d_m, const float* d_x,
                                        const float* d_logp, const float* d_grad, const float* d_step, int step_kind,
                                        int num_steps, float* d_m_out, float* d_x_out, float* d_logp_out, float* d_grad_out) {
  if (!ctx || !tgt || B < 1 || !d_m || !d_x || !d_logp || !d_grad || !d_step || num_steps < 1 || !d_m_out || !d_x_out ||
      !d_logp_out || !d_grad_out || step_kind < 0 || step_kind > 2)
    return pb2::set_error(ctx, PB2_ERR_INVALID, "pb2_logistic_tc_leapfrog: bad argument");
  if (tgt->kind != PB2_TARGET_LOGISTIC)
    return pb2::set_error(ctx, PB2_ERR_INVALID, "pb2_logistic_tc_leapfrog: target must be a logistic regression");
  cudaSetDevice(ctx->device);
  const int D = tgt->dim;
  const size_t nBD = (size_t)B * D;
  if (sizeof(float) * nBD > ctx->rs_bytes) {
    if (ctx->d_rs) cudaFree(ctx->d_rs);
    ctx->d_rs = nullptr;
    ctx->rs_bytes = 0;
    if (int rc = pb2::check_cuda(ctx, cudaMalloc((void**)&ctx->d_rs, sizeof(float) * nBD), "cudaMalloc(lock-step leapfrog scratch)"))
      return rc;
    ctx->rs_bytes = sizeof(float) * nBD;
  }
  float* v = ctx->d_rs;
  if (int rc = pb2::check_cuda(ctx, cudaMemcpyAsync(d_x_out, d_x, sizeof(float) * nBD, cudaMemcpyDeviceToDevice, ctx->stream),
                               "memcpy(x)"))
    return rc;
  if (int rc = pb2_lockstep_leapfrog(ctx, 0, B, D, d_step, step_kind, v, d_x_out, d_grad, d_m, nullptr)) return rc;
  for (int l = 0; l < num_steps; ++l) {
    if (int rc = pb2::launch_logistic_tc(ctx, const_cast<pb2_target*>(tgt), B, d_x_out, d_logp_out, d_grad_out)) return rc;
    const int last = l + 1 == num_steps;
    if (int rc = pb2_lockstep_leapfrog(ctx, last ? 2 : 1, B, D, d_step, step_kind, v, d_x_out, d_grad_out, nullptr,
                                       last ? d_m_out : nullptr))
      return rc;
  }
  (void)d_logp;
  return PB2_OK;
}

extern "C" long long pb2_rowshard_tc_planes_bytes(int N) { return N < 0 ? -1 : (long long)pb2::rowshard_tc_planes_bytes(N); }

extern "C" int pb2_rowshard_tc_prepare(pb2_ctx* ctx, const float* d_X, int N, int D, int DP, void* d_planes) {
  if (!ctx || !d_X || !d_planes || N < 0 || D < 1 || DP < D)
    return pb2::set_error(ctx, PB2_ERR_INVALID, "pb2_rowshard_tc_prepare: bad argument");
  cudaSetDevice(ctx->device);
  return pb2::launch_rowshard_tc_prepare(ctx, d_X, N, D, DP, static_cast<unsigned char*>(d_planes));
}

extern "C" int pb2_rowshard_logistic_grad_tc(pb2_ctx* ctx, const void* d_planes, const float* d_y, int N, int D,
                                             const float* d_theta, int B, float* d_packed) {
  if (!ctx || !d_planes || !d_y || !d_theta || !d_packed || N < 0 || B < 1 || D < 1)
    return pb2::set_error(ctx, PB2_ERR_INVALID, "pb2_rowshard_logistic_grad_tc: bad argument");
  cudaSetDevice(ctx->device);
  return pb2::launch_rowshard_tc(ctx, static_cast<const unsigned char*>(d_planes), d_y, N, D, d_theta, B, d_packed);
}
