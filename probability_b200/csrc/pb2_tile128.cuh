// Building blocks of the 128-chain NUTS TILE kernels (pb2_tile128_nuts.cu) on the dense-Gaussian target.
//
// The gradient of all 128 chains of a tile is ONE tcgen05 contraction G = -(X - mu) P per leapfrog (3 x 13
// tcgen05.mma M128 N112 K8, 3xTF32 split, A = (X - mu) hi/lo planes in TMEM, P hi/lo planes resident in shared
// memory, D = [128 x 112] FP32 accumulator in TMEM) -- every one of the 128 MMA rows is a chain.
//
// Thread mapping (384 threads = 8 worker warps + a third warpgroup whose first warp issues the contractions):
//   worker warp w: TMEM lane quadrant q = w & 3, half hf = w >> 2; the thread (lane l) works for chain 32 q + l
//   (= its TMEM lane) and owns the dims [52 hf, 52 hf + 52) = the four 13-dim PARTS 4 hf .. 4 hf + 3.
//   A chain is therefore TWO threads (same lane of warps q and q + 4: both may address the quadrant's TMEM lanes):
//   x, m, rho of their 52 dims live in registers; the gradient at the moving end lives in the D columns of TMEM
//   between contractions (read back per 13-dim part when needed); the previous leaf's checkpoint (m, rho) lives in
//   shared memory; everything colder in the CTA's L2 scratch.
// Synchronisation: per-chain reductions are the sums of eight 13-dim partials in a fixed tree (bit-identical in both
// threads of the pair); the two threads exchange their 4-part sums through shared memory behind a 64-thread named
// barrier (id 1 + q) -- no CTA-wide barrier inside a leaf.  The contraction is a hand-shake with the issuing warp
// through two mbarriers: the workers arrive on `full` (256 arrivals) when their A columns are staged, the issuing
// thread fires the 39 MMAs + tcgen05.commit -> `done`, the workers wait on `done`; while the tensor pipe runs they
// draw the leaf's multinomial uniform.
#pragma once
#include "pb2_tile.cuh"

namespace pb2 {
namespace tile128 {

using tile::b_plane_offset;
using tile::make_kmajor_desc;
using tile::mbar_wait;
using tile::smem_u32;
using tile::tf32_rna;
using tile::tmem_ld;
using tile::tmem_st;
using tile::tmem_wait_ld;

constexpr int kM = 128;        // chains per tile = MMA rows = TMEM lanes
constexpr int kParts = 8;      // 13-dim parts of a chain vector
constexpr int kPT = 4;         // parts per thread
constexpr int kK = 13;         // dims per part
constexpr int kDT = kPT * kK;  // dims per thread
constexpr int kKP = tile::kKP, kNP = tile::kNP;
constexpr int kWorkers = 256, kThreads = 384, kMmaWarp = 8;
constexpr int kColAhi = tile::kColAhi, kColAlo = tile::kColAlo, kColD = tile::kColD;
constexpr int kPlaneBytes = tile::kPlaneBytes;
constexpr int kRedN = 6;
constexpr size_t kVS = (size_t)kKP * kM;      // floats per scratch vector of a tile
constexpr int kPartBlk = kK * kM;             // floats per part block of a scratch vector
constexpr unsigned kFull = 0xffffffffu;

template <int V>
using ic = std::integral_constant<int, V>;

// compile-time loops: the four parts of a thread; the chunks (offset, length) that tile the 13 columns of a part
template <class F>
__device__ __forceinline__ void for_parts(F&& f) {
  f(ic<0>{}); f(ic<1>{}); f(ic<2>{}); f(ic<3>{});
}
template <class F>
__device__ __forceinline__ void for_chunks(F&& f) {
  f(ic<0>{}, ic<8>{}); f(ic<8>{}, ic<4>{}); f(ic<12>{}, ic<1>{});
}

// ---- a thread's 13-float SEGMENT of one part block of a [kKP x kM] scratch vector (global or shared memory),
// 128-bit accesses: the part block (kK * kM floats) is three [kM][4] planes followed by one [kM] plane, so a warp's
// access is contiguous.
constexpr int kSegTail = 3 * kM * 4;
template <int OFF, int N>
__device__ __forceinline__ void seg_ld(const float* pb, int cl, float* v) {
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
// one part (13 floats)
__device__ __forceinline__ void seg_ldp(const float* pb, int cl, float* v) {
  seg_ld<0, 8>(pb, cl, v); seg_ld<8, 4>(pb, cl, v + 8); seg_ld<12, 1>(pb, cl, v + 12);
}
__device__ __forceinline__ void seg_stp(float* pb, int cl, const float* v) {
  seg_st<0, 8>(pb, cl, v); seg_st<8, 4>(pb, cl, v + 8); seg_st<12, 1>(pb, cl, v + 12);
}
// all four parts of a thread; `vb` = the vector's base + the thread's first part block
__device__ __forceinline__ void seg_ldv(const float* vb, int cl, float (&v)[kDT]) {
  for_parts([&](auto pp) { constexpr int PP = decltype(pp)::value; seg_ldp(vb + PP * kPartBlk, cl, v + kK * PP); });
}
__device__ __forceinline__ void seg_stv(float* vb, int cl, const float (&v)[kDT]) {
  for_parts([&](auto pp) { constexpr int PP = decltype(pp)::value; seg_stp(vb + PP * kPartBlk, cl, v + kK * PP); });
}

struct Shared {
  unsigned long long mbar_full, mbar_done;
  uint32_t tmem_base;
  int stop;       // set before the workers' last arrival: the issuing warp leaves
  int flags[4];   // "some chain still continues" flags of the lock-step kernel, rotated per leaf
  float loc[kKP];
  float red[2][kRedN][2][kM];   // pair exchange [buffer][value][half][chain]
};

// Per-thread view of the tile.
struct Ctx {
  Shared* sh;
  uint32_t tmem, lane_addr, idesc, phase;
  uint64_t bdesc_hi, bdesc_lo;
  int cl, hf, q, parity;

  // One-time setup (all 384 threads): TMEM allocation, mbarriers, P hi/lo planes (canonical layout), loc.
  __device__ void init(Shared* sh_, unsigned char* planes, const float* P, const float* loc, int D, const float* scale) {
    sh = sh_;
    const int tid = threadIdx.x, warp = tid >> 5;
    q = warp & 3;
    hf = (warp >> 2) & 1;
    cl = 32 * q + (tid & 31);
    parity = 0;
    phase = 0;
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                   "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&sh->mbar_full)), "r"(kWorkers));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sh->mbar_done)));
      asm volatile("fence.mbarrier_init.release.cluster;");
      sh->flags[0] = sh->flags[1] = sh->flags[2] = sh->flags[3] = 0;
      sh->stop = 0;
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
    lane_addr = tmem + ((uint32_t)(32 * q) << 16);
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

  // ---- worker-side barriers
  __device__ __forceinline__ void pair_sync() const { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); }
  static __device__ __forceinline__ void wsync() { asm volatile("bar.sync 5, 256;" ::: "memory"); }
  static __device__ __forceinline__ int wsync_or(int pred) {
    int r;
    asm volatile(
        "{\n.reg .pred p, r;\nsetp.ne.b32 p, %1, 0;\nbar.red.or.pred r, 5, 256, p;\nselp.b32 %0, 1, 0, r;\n}\n"
        : "=r"(r)
        : "r"(pred)
        : "memory");
    return r;
  }

  // ---- TMEM columns of my lane: part PP of my half
  template <int PP>
  __device__ __forceinline__ void ld_cols(uint32_t col0, uint32_t (&v)[kK]) const {   // caller: tmem_wait_ld()
    const uint32_t a = lane_addr + col0 + kDT * hf + kK * PP;
    tmem_ld<8>(a, reinterpret_cast<uint32_t(&)[8]>(v[0]));
    tmem_ld<4>(a + 8, reinterpret_cast<uint32_t(&)[4]>(v[8]));
    tmem_ld<1>(a + 12, reinterpret_cast<uint32_t(&)[1]>(v[12]));
  }
  template <int PP>
  __device__ __forceinline__ void st_cols(uint32_t col0, const uint32_t (&v)[kK]) const {
    const uint32_t a = lane_addr + col0 + kDT * hf + kK * PP;
    tmem_st<8>(a, reinterpret_cast<const uint32_t(&)[8]>(v[0]));
    tmem_st<4>(a + 8, reinterpret_cast<const uint32_t(&)[4]>(v[8]));
    tmem_st<1>(a + 12, reinterpret_cast<const uint32_t(&)[1]>(v[12]));
  }
  static __device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

  // gradient at the moving end (the D columns), part PP
  template <int PP>
  __device__ __forceinline__ void ld_g(float (&g)[kK]) const {
    uint32_t t[kK];
    ld_cols<PP>(kColD, t);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < kK; ++j) g[j] = __uint_as_float(t[j]);
  }
  template <int PP>
  __device__ __forceinline__ void st_g(const float (&g)[kK]) const {
    uint32_t t[kK];
#pragma unroll
    for (int j = 0; j < kK; ++j) t[j] = __float_as_uint(g[j]);
    st_cols<PP>(kColD, t);
  }

  // xc = x - loc of part PP -> tf32 hi/lo planes of the A operand in TMEM
  template <int PP>
  __device__ __forceinline__ void stage_part(const float* xp) const {
    const float* lc = sh->loc + kDT * hf + kK * PP;
    uint32_t hi[kK], lo[kK];
#pragma unroll
    for (int j = 0; j < kK; ++j) {
      // round-to-nearest tf32 split with integer ops (cvt.rna.tf32.f32 is emulated with ~5 instructions on
      // sm_100a: inf/nan handling this path does not need -- a non-finite x gives a NaN energy = divergence)
      const float v = xp[j] - lc[j];
      hi[j] = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
      lo[j] = (__float_as_uint(v - __uint_as_float(hi[j])) + 0x1000u) & 0xffffe000u;
    }
    st_cols<PP>(kColAhi, hi);
    st_cols<PP>(kColAlo, lo);
  }

  // ---- the contraction hand-shake, worker side
  __device__ __forceinline__ void contract_begin() const {
    wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sh->mbar_full)) : "memory");
  }
  __device__ __forceinline__ void contract_end() {
    mbar_wait(smem_u32(&sh->mbar_done), phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // after the last contraction: release the issuing warp
  __device__ __forceinline__ void stop_mma() const {
    if (threadIdx.x == 0) *(volatile int*)&sh->stop = 1;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sh->mbar_full)) : "memory");
  }

  // ---- the issuing warp (24 registers after setmaxnreg.dec: the descriptors are advanced in the loop, nothing is
  // precomputed -- an unrolled sequence keeps 39 descriptors live and spills them)
  __device__ __forceinline__ void issue_pass(uint32_t a_col, uint64_t bdesc, uint32_t acc0) const {
    constexpr uint32_t step = (2u * (kNP / 8) * 128u) >> 4;               // two K core matrices per MMA (16 B units)
#pragma unroll 1
    for (int j = 0; j < kKP / 8; ++j) {
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + kColD),
          "r"(tmem + a_col + 8 * j), "l"(bdesc + (uint64_t)(j * step)), "r"(idesc), "r"(acc0 | (uint32_t)j)
          : "memory");
    }
  }
  __device__ __forceinline__ void issue_all() const {
    issue_pass(kColAhi, bdesc_hi, 0u);   // Ahi*Bhi (the first MMA overwrites D), Alo*Bhi, Ahi*Blo
    issue_pass(kColAlo, bdesc_hi, 1u);
    issue_pass(kColAhi, bdesc_lo, 1u);
  }
  __device__ void mma_loop() const {
    if ((threadIdx.x & 31) == 0) {
    uint32_t ph = 0;
    while (true) {
      mbar_wait(smem_u32(&sh->mbar_full), ph);
      ph ^= 1;
      if (*(volatile int*)&sh->stop) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_all();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sh->mbar_done))
                   : "memory");
    }
    }
    __syncwarp();
  }

  // per-chain sums: v = my four parts' sum ((p0 + p1) + (p2 + p3)); both threads of the pair return
  // lower half + upper half (identical bits); one 64-thread barrier
  template <int N>
  __device__ __forceinline__ void reduce(float (&v)[N]) {
    static_assert(N <= kRedN, "too many simultaneous reductions");
    float(*buf)[2][kM] = sh->red[parity];
    parity ^= 1;
#pragma unroll
    for (int i = 0; i < N; ++i) buf[i][hf][cl] = v[i];
    pair_sync();
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = buf[i][0][cl] + buf[i][1][cl];
  }
};

// ---- chain-major global arrays [.., B, D]: my 52 dims of chain c
__device__ __forceinline__ void tile_load(const float* base, int c, int D, int hf, bool live, float (&v)[kDT]) {
  const float* row = base + (size_t)c * D + kDT * hf;
#pragma unroll
  for (int j = 0; j < kDT; ++j) v[j] = (live && kDT * hf + j < D) ? row[j] : 0.f;
}
__device__ __forceinline__ void tile_store(float* base, size_t r, int B, int c, int D, int hf, bool live,
                                           const float (&v)[kDT]) {
  if (!live) return;
  float* row = base + (r * (size_t)B + c) * D + kDT * hf;
#pragma unroll
  for (int j = 0; j < kDT; ++j)
    if (kDT * hf + j < D) row[j] = v[j];
}

}  // namespace tile128
}  // namespace pb2
