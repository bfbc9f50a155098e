// Building blocks of the 64-chain TILE kernels (NUTS on the dense-Gaussian target).
//
// Same tensor-core contraction as pb2_tile.cuh -- G = -(X - mu) P as 3 x 13 tcgen05.mma (M128 N112 K8, 3xTF32
// split, A in TMEM, P hi/lo planes in shared memory, D in TMEM) -- but a chain occupies TWO of the 128 rows:
//   row r (TMEM lane r) = chain (r & 63), half (r >> 6); the thread (row r, slice s = warp >> 2) owns the
//   13 dims [13 part, 13 part + 13), part = 2 s + half, of its chain and writes only those columns of its row of
//   A; the other columns of the row stay zero.  So D[r] is the contribution of half the dims to the gradient and
//   g = D[r] + D[r ^ 64]; the two partner threads exchange their 13 columns through shared memory.
// Why: NUTS is bound by the latency of one leaf of ONE chain (a chain's leapfrogs are sequential and 148 SMs x 128
// rows exceed the 16,384 chains of the benchmark), and the leaf is dominated by the per-thread vector work next to
// the contraction.  13 dims per thread instead of 26 halve that work, and x, m, g, rho (52 registers) stay in
// registers without spills.  The contraction itself costs the same (the zero half-rows ride along).
// Per-chain scalars are replicated in the 8 threads of a chain and stay bit-identical because cross-part sums go
// through shared memory in a fixed order.  The 8 threads of chain c sit in rows c and c + 64, i.e. in the 8 warps with
// (warp & 1) == ((c >> 5) & 1): the half-row exchange and the cross-part reductions synchronise only those 256 threads
// (named barrier 1 + (warp & 1)), so the two halves of the tile drift apart between contractions instead of waiting
// for each other at every reduction.  (Putting both half-rows of a chain into one warp -- exchange by shuffle -- does
// not work: tcgen05.ld/st take ONE column address per warp, and the halves own different columns.)
#pragma once
#include "pb2_tile.cuh"

namespace pb2 {
namespace tile64 {

using tile::b_plane_offset;
using tile::make_kmajor_desc;
using tile::mbar_wait;
using tile::smem_u32;
using tile::tf32_rna;
using tile::tmem_ld;
using tile::tmem_st;
using tile::tmem_wait_ld;

constexpr int kRows = 128;    // MMA M = TMEM lanes
constexpr int kM = 64;        // chains per tile
constexpr int kSlices = 4;
constexpr int kParts = 8;     // (slice, half)
constexpr int kK = 13;        // dims per thread
constexpr int kKP = tile::kKP, kNP = tile::kNP, kThreads = tile::kThreads;
constexpr int kColAhi = tile::kColAhi, kColAlo = tile::kColAlo, kColD = tile::kColD;
constexpr int kPlaneBytes = tile::kPlaneBytes;
constexpr int kRedN = 6;
constexpr size_t kVS = (size_t)kKP * kM;   // floats per scratch vector of a tile

// compile-time loop over the chunks (offset, length) that tile the 13 columns of a part
template <class F>
__device__ __forceinline__ void for_chunks(F&& f) {
  f(std::integral_constant<int, 0>{}, std::integral_constant<int, 8>{});
  f(std::integral_constant<int, 8>{}, std::integral_constant<int, 4>{});
  f(std::integral_constant<int, 12>{}, std::integral_constant<int, 1>{});
}

// ---- a thread's 13-float SEGMENT of a [kKP x kM] scratch vector (global or shared memory), 128-bit accesses:
// the part block (kK * kM floats) is three [kM][4] planes followed by one [kM] plane.
constexpr int kSegTail = 3 * kM * 4;
template <int OFF, int N>
__device__ __forceinline__ void seg_ld(const float* pb, int cl, float (&v)[N]) {
  if constexpr (N == 1) {
    v[0] = pb[kSegTail + cl];
  } else {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(pb + ((OFF / 4 + q) * kM + cl) * 4);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  }
}
template <int OFF, int N>
__device__ __forceinline__ void seg_st(float* pb, int cl, const float* v) {
  if constexpr (N == 1) {
    pb[kSegTail + cl] = v[0];
  } else {
#pragma unroll
    for (int q = 0; q < N / 4; ++q)
      *reinterpret_cast<float4*>(pb + ((OFF / 4 + q) * kM + cl) * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
}
__device__ __forceinline__ void seg_ldv(const float* pb, int cl, float (&v)[kK]) {
  float a[8], b[4], c[1];
  seg_ld<0, 8>(pb, cl, a); seg_ld<8, 4>(pb, cl, b); seg_ld<12, 1>(pb, cl, c);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = a[j];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[8 + j] = b[j];
  v[12] = c[0];
}
__device__ __forceinline__ void seg_stv(float* pb, int cl, const float (&v)[kK]) {
  seg_st<0, 8>(pb, cl, v); seg_st<8, 4>(pb, cl, v + 8); seg_st<12, 1>(pb, cl, v + 12);
}

struct Shared {
  unsigned long long mbar;
  uint32_t tmem_base;
  int flags[4];   // "some chain still continues" flags of the lock-step kernel, rotated per leaf
  float loc[kKP];
  float red[2][kRedN][kParts][kM];
};
constexpr int kXbufFloats = kParts * 4 * kM * 4;   // gradient exchange: [part][plane 0..3][chain][4]

// Per-thread view of the tile.
struct Ctx {
  Shared* sh;
  float* xbuf;
  uint32_t tmem, lane_addr, idesc, phase;
  uint64_t bdesc_hi, bdesc_lo;
  int cl, slice, half, part, parity, group;

  // One-time setup: TMEM allocation, mbarrier, P hi/lo planes (canonical layout), loc, zero A.
  __device__ void init(Shared* sh_, unsigned char* planes, float* xbuf_, const float* P, const float* loc, int D, const float* scale = nullptr) {
    sh = sh_;
    xbuf = xbuf_;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = 32 * (warp & 3) + (tid & 31);
    cl = row & (kM - 1);
    half = row >> 6;
    slice = warp >> 2;
    part = 2 * slice + half;
    group = warp & 1;
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
    idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((uint32_t)(kNP >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
    // the columns of my row that belong to the other half of my chain are never written again: zero them
    {
      const uint32_t base = lane_addr + kK * (part ^ 1);
      for_chunks([&](auto off, auto n) {
        constexpr int OFF = decltype(off)::value, N = decltype(n)::value;
        uint32_t z[N];
#pragma unroll
        for (int j = 0; j < N; ++j) z[j] = 0u;
        tmem_st<N>(base + kColAhi + OFF, z);
        tmem_st<N>(base + kColAlo + OFF, z);
      });
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }

  __device__ void finish() {
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }

  // xc = x - loc of my part -> tf32 hi/lo planes of the A operand in TMEM
  __device__ __forceinline__ void stage_a(const float (&x)[kK]) {
    const uint32_t base = lane_addr + kK * part;
    const float* lc = sh->loc + kK * part;
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

  // all threads: A is staged -> thread 0 issues 3 x 13 MMAs + commit -> everybody waits -> the two half-rows of a
  // chain exchange their contributions -> g = gradient of my 13 dims
  __device__ __forceinline__ void contract(float (&g)[kK]) {
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
    // my row of D: my own 13 columns and the 13 columns my partner thread (row ^ 64, same slice) owns
    const uint32_t mine = lane_addr + kColD + kK * part, theirs = lane_addr + kColD + kK * (part ^ 1);
    uint32_t a0[8], a1[4], a2[1], b0[8], b1[4], b2[1];
    tmem_ld<8>(theirs, b0); tmem_ld<4>(theirs + 8, b1); tmem_ld<1>(theirs + 12, b2);
    tmem_ld<8>(mine, a0); tmem_ld<4>(mine + 8, a1); tmem_ld<1>(mine + 12, a2);
    tmem_wait_ld();
    float4* xw = reinterpret_cast<float4*>(xbuf) + (size_t)(part ^ 1) * 4 * kM + cl;
    xw[0] = make_float4(__uint_as_float(b0[0]), __uint_as_float(b0[1]), __uint_as_float(b0[2]), __uint_as_float(b0[3]));
    xw[kM] = make_float4(__uint_as_float(b0[4]), __uint_as_float(b0[5]), __uint_as_float(b0[6]), __uint_as_float(b0[7]));
    xw[2 * kM] = make_float4(__uint_as_float(b1[0]), __uint_as_float(b1[1]), __uint_as_float(b1[2]), __uint_as_float(b1[3]));
    reinterpret_cast<float*>(xw + 3 * kM)[0] = __uint_as_float(b2[0]);
    group_sync();
    const float4* xr = reinterpret_cast<const float4*>(xbuf) + (size_t)part * 4 * kM + cl;
    const float4 r0 = xr[0], r1 = xr[kM], r2 = xr[2 * kM];
    const float r3 = reinterpret_cast<const float*>(xr + 3 * kM)[0];
    g[0] = __uint_as_float(a0[0]) + r0.x; g[1] = __uint_as_float(a0[1]) + r0.y;
    g[2] = __uint_as_float(a0[2]) + r0.z; g[3] = __uint_as_float(a0[3]) + r0.w;
    g[4] = __uint_as_float(a0[4]) + r1.x; g[5] = __uint_as_float(a0[5]) + r1.y;
    g[6] = __uint_as_float(a0[6]) + r1.z; g[7] = __uint_as_float(a0[7]) + r1.w;
    g[8] = __uint_as_float(a1[0]) + r2.x; g[9] = __uint_as_float(a1[1]) + r2.y;
    g[10] = __uint_as_float(a1[2]) + r2.z; g[11] = __uint_as_float(a1[3]) + r2.w;
    g[12] = __uint_as_float(a2[0]) + r3;
  }

  // the 256 threads that own chains with ((c >> 5) & 1) == group
  __device__ __forceinline__ void group_sync() const { asm volatile("bar.sync %0, 256;" ::"r"(1 + group) : "memory"); }

  // cross-part sums (fixed order => the 8 threads of a chain get identical bits); one group barrier
  template <int N>
  __device__ __forceinline__ void reduce(float (&v)[N]) {
    static_assert(N <= kRedN, "too many simultaneous reductions");
    float(*buf)[kParts][kM] = sh->red[parity];
    parity ^= 1;
#pragma unroll
    for (int i = 0; i < N; ++i) buf[i][part][cl] = v[i];
    group_sync();
#pragma unroll
    for (int i = 0; i < N; ++i)
      v[i] = (((buf[i][0][cl] + buf[i][1][cl]) + (buf[i][2][cl] + buf[i][3][cl])) +
              ((buf[i][4][cl] + buf[i][5][cl]) + (buf[i][6][cl] + buf[i][7][cl])));
  }
};

// ---- chain-major global arrays [.., B, D]: my 13 dims of chain c
__device__ __forceinline__ void tile_load(const float* base, int c, int D, int part, bool live, float (&v)[kK]) {
  const float* row = base + (size_t)c * D + kK * part;
#pragma unroll
  for (int j = 0; j < kK; ++j) v[j] = (live && kK * part + j < D) ? row[j] : 0.f;
}
__device__ __forceinline__ void tile_store(float* base, size_t r, int B, int c, int D, int part, bool live,
                                           const float (&v)[kK]) {
  if (!live) return;
  float* row = base + (r * (size_t)B + c) * D + kK * part;
#pragma unroll
  for (int j = 0; j < kK; ++j)
    if (kK * part + j < D) row[j] = v[j];
}

}  // namespace tile64
}  // namespace pb2
