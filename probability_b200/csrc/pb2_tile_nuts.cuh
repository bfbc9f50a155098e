// One NUTS subtree (= one tree doubling, nuts.py:713-946 `_build_sub_tree` / `_loop_build_sub_tree`) for a
// 128-chain tile in lock-step; shared by tile_nuts_kernel and tile_nuts_sched_kernel.
//
// Where the state of my (chain, 26-dim slice) lives during a subtree:
//   registers     : x, m (moving end), rho (cumulative momentum of the subtree)
//   TMEM D        : g = gradient at x -- the accumulator of the last contraction is read back chunk by
//                   chunk where it is consumed, it never occupies 26 registers across a leaf
//   shared memory : the checkpoint written by the previous (even) leaf -- every odd leaf checks against it
//   L2 scratch    : the popcount-indexed checkpoint stores that later leaves need (only leaves with
//                   i % 4 == 0 are ever read again after the next leaf) and the subtree candidate
// Every per-thread segment uses the 128-bit layout of pb2_tile.cuh (seg_ld / seg_st).
#pragma once
#include "pb2_tile.cuh"

namespace pb2 {
namespace tile {

#ifdef PB2_TILE_PROF
__device__ unsigned long long g_tile_prof[2][16];
struct Prof {
  int w;
  long long t;
  __device__ void init() {
    w = (blockIdx.x == 0 && threadIdx.x == 0) ? 0 : ((blockIdx.x == 0 && threadIdx.x == kThreads - 1) ? 1 : -1);
    t = clock64();
  }
  __device__ __forceinline__ void mark(int k) {
    if (w >= 0) {
      const long long now = clock64();
      atomicAdd(&g_tile_prof[w][k], (unsigned long long)(now - t));
      t = now;
    }
  }
  __device__ __forceinline__ void leaf() { if (w >= 0) atomicAdd(&g_tile_prof[w][15], 1ull); }
};
#else
struct Prof {
  __device__ void init() {}
  __device__ __forceinline__ void mark(int) {}
  __device__ __forceinline__ void leaf() {}
};
#endif

struct SubtreeState {
  float slp;        // in: log-prob at x; out: log-prob at the new end
  float blp, ben;   // subtree candidate's log-prob and energy
  float bw;         // log-sum of the subtree's weights
  float esum_sub;   // sum of min(1, exp(dH)) over the subtree's leaves (for log_accept_ratio)
  int n;            // leaves this chain took
  bool c_prev;      // in: chain continues; out: subtree finished without U-turn / divergence
  bool nd;          // not diverged
  bool took;        // out: a new subtree candidate was written to the scratch (bx, bg) during this call
};

struct SubtreeArgs {
  const uint32_t* kud;   // multinomial keys of this depth: 2 words per leaf
  int nsteps;            // 2^depth leaves
  int unrolled;          // unrolled_leapfrog_steps
  int layout;            // counter layout of the bit generator
  uint64_t cg, b_global; // global chain index / number of chains (RNG counters)
  float eps, H0, lognorm, max_energy_diff;
  const float* lc;       // loc of my slice (shared memory)
  float* bx;             // L2 scratch, slice bases: subtree candidate (x, g)
  float* bg;
  float* ck_m;           // checkpoint stores: momentum of slot k at ck_m + k * kVS, rho at ck_r + k * kVS
  float* ck_r;
  float* ckl;            // shared memory, slice base: last even leaf's checkpoint (momentum; rho at + kVS)
  // kMixed only (32-leaf chunks of doublings >= 5, every chain at its own chunk of its own doubling):
  float* hck;            // this chain's record: checkpoints of chunk-first leaves, slot s = (m at 2s, rho at 2s + 1) * kRecStride
  int ihi;               // my chunk index within my doubling
  int hi_slot_w;         // >= 0: leaf 0 of this chunk is checked again by later chunks -> store it in slot hi_slot_w
  int hi_checks;         // tile-wide max of trailing_ones(ihi): number of cross-chunk U-turn checks at the last leaf
};

// ---- chain-major records (pb2_tile_sched.cuh): a vector is 4 slices x 28 floats (26 used; 16-byte aligned segments)
constexpr int kRecSlice = 28, kRecStride = 4 * kRecSlice;
template <int OFF, int N>
__device__ __forceinline__ void rec_ld(const float* vb, float (&v)[N]) {
  if constexpr (N == 2) {
    const float2 t = __ldcg(reinterpret_cast<const float2*>(vb + OFF));
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 t = __ldcg(reinterpret_cast<const float4*>(vb + OFF + 4 * q));
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  }
}
template <int OFF, int N>
__device__ __forceinline__ void rec_st(float* vb, const float* v) {
  if constexpr (N == 2) {
    __stcg(reinterpret_cast<float2*>(vb + OFF), make_float2(v[0], v[1]));
  } else {
#pragma unroll
    for (int q = 0; q < N / 4; ++q)
      __stcg(reinterpret_cast<float4*>(vb + OFF + 4 * q), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
  }
}
__device__ __forceinline__ void rec_ld26(const float* vb, float (&v)[kK]) {
  float a[16], b[8], c[2];
  rec_ld<0, 16>(vb, a); rec_ld<16, 8>(vb, b); rec_ld<24, 2>(vb, c);
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = a[j];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[16 + j] = b[j];
  v[24] = c[0]; v[25] = c[1];
}
__device__ __forceinline__ void rec_st26(float* vb, const float (&v)[kK]) {
  rec_st<0, 16>(vb, v); rec_st<16, 8>(vb, v + 16); rec_st<24, 2>(vb, v + 24);
}

constexpr size_t kVS = (size_t)kKP * kM;   // floats per scratch vector

// 2 U-turn checks against stored checkpoints (nuts.py:949-1010): s[2q] = <rho - rho_k, m_k>, s[2q+1] = <rho - rho_k, m>
__device__ __forceinline__ void uturn_pair(const float* km0, const float* kr0, const float* km1, const float* kr1, int cl,
                                           const float (&rho)[kK], const float (&m)[kK], float (&s)[4]) {
  for_chunks([&](auto off, auto nn) {
    constexpr int OFF = decltype(off)::value, N = decltype(nn)::value;
    constexpr int W = N == 2 ? 2 : 4;
#pragma unroll
    for (int q = 0; q < N / W; ++q) {
      float a0[W], b0[W], a1[W], b1[W];
      if constexpr (W == 2) {
        seg_ld<24, 2>(kr0, cl, a0); seg_ld<24, 2>(km0, cl, b0); seg_ld<24, 2>(kr1, cl, a1); seg_ld<24, 2>(km1, cl, b1);
      } else {
        // one float4 of each of the four vectors (OFF + 4q .. +3)
        const int o = ((OFF / 4 + q) * kM + cl) * 4;
        const float4 ta0 = *reinterpret_cast<const float4*>(kr0 + o), tb0 = *reinterpret_cast<const float4*>(km0 + o);
        const float4 ta1 = *reinterpret_cast<const float4*>(kr1 + o), tb1 = *reinterpret_cast<const float4*>(km1 + o);
        a0[0] = ta0.x; a0[1] = ta0.y; a0[2] = ta0.z; a0[3] = ta0.w;
        b0[0] = tb0.x; b0[1] = tb0.y; b0[2] = tb0.z; b0[3] = tb0.w;
        a1[0] = ta1.x; a1[1] = ta1.y; a1[2] = ta1.z; a1[3] = ta1.w;
        b1[0] = tb1.x; b1[1] = tb1.y; b1[2] = tb1.z; b1[3] = tb1.w;
      }
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const int d = OFF + W * q + j;
        const float d0 = rho[d] - a0[j], d1 = rho[d] - a1[j];
        s[0] = fmaf(d0, b0[j], s[0]);
        s[1] = fmaf(d0, m[d], s[1]);
        s[2] = fmaf(d1, b1[j], s[2]);
        s[3] = fmaf(d1, m[d], s[3]);
      }
    }
  });
}

// On entry: x, m = the end that is extended, g(x) in TMEM D.  On exit the same for the new end, rho = sum of the
// subtree's momenta, st = the subtree's scalars.  `gt` counts leaves globally (it rotates the flags through which
// the tile learns that no chain continues, nuts.py:759 reduce_any(continue_tree)).
template <bool kMixed>
__device__ __forceinline__ void nuts_subtree(Ctx& cx, Shared& sh, float (*lu)[kM], unsigned& gt, const SubtreeArgs& a,
                                             float (&x)[kK], float (&m)[kK], float (&rho)[kK], SubtreeState& st, Prof& pf) {
  const float eps = a.eps, heps = 0.5f * a.eps;
  const int cl = cx.cl;
  float slp = st.slp, blp = st.slp, ben = st.slp, bw = -INFINITY, esum_sub = 0.f;
  int n = 0;
  bool c_prev = st.c_prev, nd = st.nd, took = false;
  if constexpr (kMixed) {   // a chunk continues a subtree: the caller loaded (or initialised) rho and the scalars
    blp = st.blp; ben = st.ben; bw = st.bw; esum_sub = st.esum_sub; n = st.n;
  } else {
#pragma unroll
    for (int j = 0; j < kK; ++j) rho[j] = 0.f;
  }
#pragma unroll 1
  for (int i = 0; i < a.nsteps; ++i, ++gt) {
    pf.mark(0);
    if ((i & 3) == 0 && i + cx.slice < a.nsteps) {   // 4 leaves of multinomial uniforms, one per slice
      Key kk{a.kud[2 * (i + cx.slice)], a.kud[2 * (i + cx.slice) + 1]};
      lu[cx.slice][cl] = log1pf(-uniform_from_bits(bits_at(kk, a.cg, a.b_global, a.layout), 0.f, 1.f));
    }
    // one leapfrog (leapfrog_integrator.py:280-309 with L = unrolled_leapfrog_steps): half kick from the g in D
    for_chunks([&](auto off, auto nn) {
      constexpr int OFF = decltype(off)::value, N = decltype(nn)::value;
      float gc[N];
      cx.load_d_chunk<OFF, N>(gc);
#pragma unroll
      for (int j = 0; j < N; ++j) m[OFF + j] = m[OFF + j] + heps * gc[j];
    });
    bool stop = false;
    float lu_i = 0.f;
#pragma unroll 1
    for (int l = 0; l < a.unrolled; ++l) {
#pragma unroll
      for (int j = 0; j < kK; ++j) x[j] = x[j] + eps * m[j];
      cx.stage_a(x);
      pf.mark(1);
      cx.contract();
      pf.mark(2);
      if (l == 0) {
        // read before the next barrier: a faster slice may overwrite lu[] for the next 4 leaves after it
        lu_i = lu[i & 3][cl];
        // flag raised at the end of the previous leaf; leaving mid-leaf is harmless because no chain of the
        // tile continues (ends / candidates are final)
        if (i > 0 && sh.flags[(gt - 1) & 3] == 0) stop = true;
        if (threadIdx.x == 0) sh.flags[(gt + 1) & 3] = 0;
      }
      if (stop) break;
      if (l + 1 < a.unrolled) {
        for_chunks([&](auto off, auto nn) {
          constexpr int OFF = decltype(off)::value, N = decltype(nn)::value;
          float gc[N];
          cx.load_d_chunk<OFF, N>(gc);
#pragma unroll
          for (int j = 0; j < N; ++j) m[OFF + j] = m[OFF + j] + eps * gc[j];
        });
      }
    }
    if (stop) break;
    n += c_prev ? 1 : 0;
    // last kick, rho_subtree, checkpoint store / first U-turn check (nuts.py:826-869, 949-1010), chunk by chunk
    float s4[4] = {0.f, 0.f, 0.f, 0.f};   // <x - mu, g>, |m|^2, U-turn dots against the previous leaf's checkpoint
    const int pc = __popc(i);
    const bool odd = (i & 1) != 0;
    const int ones = __ffs(~i) - 1;          // trailing ones: the leaf closes subtrees of 2, 4, .., 2^ones leaves
    const bool keep = (i & 3) == 0;          // an even leaf that is checked again after leaf i + 1
    float* const ckm_w = a.ck_m + (size_t)pc * kVS;
    float* const ckr_w = a.ck_r + (size_t)pc * kVS;
    for_chunks([&](auto off, auto nn) {
      constexpr int OFF = decltype(off)::value, N = decltype(nn)::value;
      float gc[N];
      cx.load_d_chunk<OFF, N>(gc);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        m[OFF + j] = m[OFF + j] + eps * gc[j];
        m[OFF + j] = m[OFF + j] - heps * gc[j];
        s4[0] = fmaf(x[OFF + j] - a.lc[OFF + j], gc[j], s4[0]);
        s4[1] = fmaf(m[OFF + j], m[OFF + j], s4[1]);
      }
      if (!odd) {
        seg_st<OFF, N>(a.ckl, cl, m + OFF);
        seg_st<OFF, N>(a.ckl + kVS, cl, rho + OFF);
        if (keep) {
          seg_st<OFF, N>(ckm_w, cl, m + OFF);
          seg_st<OFF, N>(ckr_w, cl, rho + OFF);
        }
        if constexpr (kMixed) {
          if (i == 0 && a.hi_slot_w >= 0) {
            rec_st<OFF, N>(a.hck + (size_t)(2 * a.hi_slot_w) * kRecStride, m + OFF);
            rec_st<OFF, N>(a.hck + (size_t)(2 * a.hi_slot_w + 1) * kRecStride, rho + OFF);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < N; ++j) rho[OFF + j] = rho[OFF + j] + m[OFF + j];
      if (odd) {
        float km[N], kr[N];
        seg_ld<OFF, N>(a.ckl, cl, km);
        seg_ld<OFF, N>(a.ckl + kVS, cl, kr);
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float diff = rho[OFF + j] - kr[j];
          s4[2] = fmaf(diff, km[j], s4[2]);
          s4[3] = fmaf(diff, m[OFF + j], s4[3]);
        }
      }
    });
    pf.mark(3);
    cx.reduce<4>(s4);
    pf.mark(4);
    slp = fmaf(0.5f, s4[0], a.lognorm);
    bool ok = true;
    if (odd) {
      ok = (s4[2] >= 0.f) && (s4[3] >= 0.f);
      // the larger subtrees this leaf closes: slots [pc - ones, pc - 1), two checks per reduction
#pragma unroll 1
      for (int k = pc - ones; k < pc - 1; k += 2) {   // uniform trip count over the tile (lock-step leaf index)
        const bool two = k + 1 < pc - 1;
        const int k1 = two ? k + 1 : k;
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        uturn_pair(a.ck_m + (size_t)k * kVS, a.ck_r + (size_t)k * kVS, a.ck_m + (size_t)k1 * kVS,
                   a.ck_r + (size_t)k1 * kVS, cl, rho, m, s);
        cx.reduce<4>(s);
        ok = ok && (s[0] >= 0.f) && (s[1] >= 0.f) && (s[2] >= 0.f) && (s[3] >= 0.f);
      }
      if constexpr (kMixed) {
        // last leaf of the chunk: the subtrees of 64, 128, .. leaves it closes start at the first leaf of an
        // earlier chunk of this chain's doubling (per-chain slots in the chain's record)
        if (i == a.nsteps - 1) {
          const int t_hi = __ffs(~a.ihi) - 1;
#pragma unroll 1
          for (int jj = 1; jj <= a.hi_checks; ++jj) {
            const bool act = jj <= t_hi;
            const int sl = __popc(a.ihi - (1 << jj) + 1);
            const float* km = a.hck + (size_t)(2 * sl) * kRecStride;
            const float* kr = km + kRecStride;
            float s2[2] = {0.f, 0.f};
            if (act) {
              for_chunks([&](auto off, auto nn) {
                constexpr int OFF = decltype(off)::value, N = decltype(nn)::value;
                float vm[N], vr[N];
                rec_ld<OFF, N>(km, vm);
                rec_ld<OFF, N>(kr, vr);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                  const float diff = rho[OFF + j] - vr[j];
                  s2[0] = fmaf(diff, vm[j], s2[0]);
                  s2[1] = fmaf(diff, m[OFF + j], s2[1]);
                }
              });
            }
            cx.reduce<2>(s2);
            ok = ok && (!act || ((s2[0] >= 0.f) && (s2[1] >= 0.f)));
          }
        }
      }
    }
    pf.mark(5);
    float en = slp - 0.5f * s4[1];                 // nuts.py:871-877
    en = isnan(en) ? -INFINITY : en;
    const float dH = en - a.H0;
    const bool nd_i = (-dH) < a.max_energy_diff;   // :880
    const float w_new = log_add_exp(bw, dH);       // :881-883
    const bool take = lu_i <= (dH - w_new);        // :897-901
    if (__any_sync(0xffffffffu, take)) {
      float g[kK];
      cx.load_d(g);
      if (take) {
        seg_st26(a.bx, cl, x);
        seg_st26(a.bg, cl, g);
        blp = slp; ben = en;
        took = true;
      }
    }
    bw = w_new;
    const bool c_now = nd_i && c_prev;             // :921
    if (c_now) esum_sub += expf(fminf(dH, 0.f));   // :930-933
    nd = nd && (c_prev ? nd_i : true);             // :924-927,944
    c_prev = ok && c_now;                          // :922
    if (c_prev) sh.flags[gt & 3] = 1;
    pf.mark(6);
    pf.leaf();
  }
  st.slp = slp; st.blp = blp; st.ben = ben; st.bw = bw; st.esum_sub = esum_sub;
  st.n = n; st.c_prev = c_prev; st.nd = nd; st.took = took;
}

}  // namespace tile
}  // namespace pb2
