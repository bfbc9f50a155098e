// 64-chain TILE NUTS kernels for the dense-Gaussian target (tcgen05 path, pb2_tile64.cuh):
//   tile_nuts_kernel       : NoUTurnSampler.one_step (tfp/mcmc/nuts.py:321-946) for a tile run in LOCK-STEP, i.e.
//                            literally the reference's batched algorithm (shared doubling / leaf counters,
//                            per-chain masks) -- max_tree_depth <= 5, or dense_variant 3 (the bit-exact partner of
//                            the asynchronous kernel in the tests);
//   tile_nuts_async_kernel : every lane at its own position of its own tree and transition -- all other launches
//                            (single transitions, adaptation steps and fused multi-transition runs).
// Both call the same nuts_leaf(), so a chain's arithmetic is the same instruction sequence in both and the
// results are bit-identical (tests/test_gpu_parity.py).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "pb2_tile64.cuh"

namespace pb2 {
using namespace tile64;

constexpr int kColS4 = kColD + kNP;   // TMEM columns 320 .. 423: the 4-leaf-subtree checkpoint (m | rho per slice)
static_assert(kColS4 + kSlices * 2 * kK <= 512, "TMEM budget");

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

__device__ __forceinline__ int nuts_result_index(const ChainParams& p, int t) {
  int u = t - p.burnin;
  if (u < 0) return -1;
  int q = u / (p.thin + 1);
  if (q * (p.thin + 1) != u || q >= p.n_results) return -1;
  return q;
}

// momentum ~ N(0, I): one key per state part, counter = row-major index in [B_global, size_part]
__device__ __forceinline__ float nuts_momentum(const ChainParams& p, const uint32_t* keys, uint64_t cg, int d) {
  int part = 0;
#pragma unroll 1
  for (int q = 1; q < p.n_parts; ++q) part += (d >= p.part_off[q]) ? 1 : 0;
  const int off = p.part_off[part];
  const uint64_t sz = (uint64_t)(p.part_off[part + 1] - off);
  Key k{keys[2 * part], keys[2 * part + 1]};
  return normal_from_bits(bits_at(k, cg * sz + (uint64_t)(d - off), (uint64_t)p.B_global * sz, p.layout));
}

// ---------------------------------------------------------------------------------------------
// One leaf of a NUTS subtree (nuts.py:793-946 `_loop_build_sub_tree` body) for every lane of the tile.
// The tile shares the leaf clock `i` (parity, popcount slot, which 2-, 4-, .. leaf subtrees close); everything
// else is per lane:
//   act   : the lane takes this leaf (it continues its subtree, and it is not an idle tick of its schedule)
//   jmax  : the largest subtree (2^jmax leaves) that can close inside the lane's doubling at this leaf
//   hi_*  : (async kernel) checkpoints of the first leaves of the lane's earlier 32-leaf chunks
// Where my 13 dims live: registers x, m, g (moving end), rho (cumulative momentum of the subtree); shared memory
// ckl = the checkpoint written by the previous (even) leaf; L2 scratch = the popcount-indexed checkpoint slots that
// later leaves need (only leaves with i % 4 == 0 are read again after leaf i + 1) and the subtree candidate.
struct LaneSub {
  float slp;        // log-prob at the moving end
  float blp, ben;   // subtree candidate's log-prob and energy
  float bw;         // log-sum of the subtree's weights
  float esum_sub;   // sum of min(1, exp(dH)) over the subtree's leaves (for log_accept_ratio)
  int n;            // leaves taken
  bool alive;       // no U-turn / divergence inside the subtree so far
  bool nd;          // not diverged
};

struct LeafEnv {
  const float* lc;        // loc of my part (shared memory)
  float* bx;              // part bases: subtree candidate (x, g)
  float* bg;
  float* ck_m;            // checkpoint slot k: momentum at ck_m + k * kVS, rho at ck_r + k * kVS
  float* ck_r;
  float* ckl;             // shared memory: previous even leaf's checkpoint (momentum; rho at + kVS)
  float* hi_m;            // per-lane slots of chunk-first leaves (async kernel)
  float* hi_r;
  int* flags;             // lock-step kernel: "some chain continues" flags (nullptr in the async kernel)
  int unrolled;
  float lognorm, max_energy_diff;
};

// 2 U-turn checks against stored checkpoints (nuts.py:949-1010): s[2q] = <rho - rho_k, m_k>, s[2q+1] = <rho - rho_k, m>
__device__ __forceinline__ void uturn_pair(const float* km0, const float* kr0, const float* km1, const float* kr1, int cl,
                                           const float (&rho)[kK], const float (&m)[kK], float (&s)[4]) {
  for_chunks([&](auto off, auto nn) {
    constexpr int OFF = decltype(off)::value, N = decltype(nn)::value;
    float a0[N], b0[N], a1[N], b1[N];
    seg_ld<OFF, N>(kr0, cl, a0); seg_ld<OFF, N>(km0, cl, b0);
    seg_ld<OFF, N>(kr1, cl, a1); seg_ld<OFF, N>(km1, cl, b1);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float d0 = rho[OFF + j] - a0[j], d1 = rho[OFF + j] - a1[j];
      s[0] = fmaf(d0, b0[j], s[0]);
      s[1] = fmaf(d0, m[OFF + j], s[1]);
      s[2] = fmaf(d1, b1[j], s[2]);
      s[3] = fmaf(d1, m[OFF + j], s[3]);
    }
  });
}

// returns false if the lock-step tile stopped (no chain continues)
__device__ __forceinline__ bool nuts_leaf(Ctx& cx, const LeafEnv& e, int i, unsigned gt, bool act, int jmax, int hi_slot_w,
                                          int ihi, int t_hi, int hi_checks, float eps, float H0, const float* lu_row,
                                          float (&x)[kK], float (&m)[kK], float (&g)[kK], float (&rho)[kK], LaneSub& s,
                                          Prof& pf) {
  const int cl = cx.cl;
  const float heps = 0.5f * eps;
  // one leapfrog (leapfrog_integrator.py:280-309 with L = unrolled_leapfrog_steps)
  if (act) {
#pragma unroll
    for (int j = 0; j < kK; ++j) m[j] = m[j] + heps * g[j];
  }
  float lu_i = 0.f;
#pragma unroll 1
  for (int l = 0; l < e.unrolled; ++l) {
    if (act) {
#pragma unroll
      for (int j = 0; j < kK; ++j) x[j] = x[j] + eps * m[j];
    }
    cx.stage_a(x);
    pf.mark(1);
    float gn[kK];
    cx.contract(gn);
    pf.mark(2);
    if (l == 0) {
      lu_i = lu_row[cl];   // read before the next barrier: the multinomial draws are rewritten every 4 leaves
      if (e.flags) {
        // flag raised at the end of the previous leaf; leaving mid-leaf is harmless because no chain of the
        // tile continues (ends / candidates are final)
        if (i > 0 && e.flags[(gt - 1) & 3] == 0) return false;
        if (threadIdx.x == 0) e.flags[(gt + 1) & 3] = 0;
      }
    }
    if (act) {
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        g[j] = gn[j];
        m[j] = m[j] + eps * g[j];
      }
    }
  }
  // last half kick, rho_subtree, checkpoint store / U-turn checks of the closing 2- and 4-leaf subtrees
  // (nuts.py:826-869, 949-1010)
  float s6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // <x - mu, g>, |m|^2, dots vs the previous leaf, dots vs slot pc - 2
  const int pc = __popc(i);
  const bool odd = (i & 1) != 0;
  const int ones = __ffs(~i) - 1;          // trailing ones: the leaf closes subtrees of 2, 4, .., 2^ones leaves
  const bool slot1 = odd && ones >= 2;     // (tile-uniform) the 4-leaf subtree closes: its checkpoint is slot pc - 2
  // The checkpoint of a 4-leaf subtree's first leaf (i % 4 == 0) is read back three leaves later: it lives in the free
  // TMEM columns of the thread's own lane (26 columns per slice: m | rho) instead of the L2 scratch, which only keeps the
  // first leaves of the 8-, 16-, 32-leaf subtrees (i % 8 == 0).  TMEM accesses are warp-collective: they sit outside the
  // per-lane `act` branches (an inactive lane parks garbage in its own columns).
  const bool s4_store = !odd && (i & 3) == 0;
  const uint32_t s4_addr = cx.lane_addr + kColS4 + 2 * kK * cx.slice;
  if (act) {
#pragma unroll
    for (int j = 0; j < kK; ++j) {
      m[j] = m[j] - heps * g[j];
      s6[0] = fmaf(x[j] - e.lc[j], g[j], s6[0]);
      s6[1] = fmaf(m[j], m[j], s6[1]);
    }
    if (!odd) {
      seg_stv(e.ckl, cl, m);
      seg_stv(e.ckl + kVS, cl, rho);
      if ((i & 7) == 0) {   // the first leaf of an 8-leaf (or larger) subtree: checked again after leaf i + 3
        seg_stv(e.ck_m + (size_t)pc * kVS, cl, m);
        seg_stv(e.ck_r + (size_t)pc * kVS, cl, rho);
      }
      if (i == 0 && hi_slot_w >= 0) {
        seg_stv(e.hi_m + (size_t)hi_slot_w * kVS, cl, m);
        seg_stv(e.hi_r + (size_t)hi_slot_w * kVS, cl, rho);
      }
    }
  }
  auto s4_st = [&](uint32_t a, const float (&v)[kK]) {   // 13 columns straight from the registers
    tmem_st<8>(a, reinterpret_cast<const uint32_t(&)[8]>(v[0]));
    tmem_st<4>(a + 8, reinterpret_cast<const uint32_t(&)[4]>(v[8]));
    tmem_st<1>(a + 12, reinterpret_cast<const uint32_t(&)[1]>(v[12]));
  };
  auto s4_ld = [&](uint32_t a, float (&v)[kK]) {
    tmem_ld<8>(a, reinterpret_cast<uint32_t(&)[8]>(v[0]));
    tmem_ld<4>(a + 8, reinterpret_cast<uint32_t(&)[4]>(v[8]));
    tmem_ld<1>(a + 12, reinterpret_cast<uint32_t(&)[1]>(v[12]));
  };
  if (s4_store) {
    s4_st(s4_addr, m);
    s4_st(s4_addr + kK, rho);
  }
  if (act) {
#pragma unroll
    for (int j = 0; j < kK; ++j) rho[j] = rho[j] + m[j];
    if (odd) {
      float km[kK], kr[kK];
      seg_ldv(e.ckl, cl, km);
      seg_ldv(e.ckl + kVS, cl, kr);
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        const float diff = rho[j] - kr[j];
        s6[2] = fmaf(diff, km[j], s6[2]);
        s6[3] = fmaf(diff, m[j], s6[3]);
      }
    }
  }
  if (slot1) {   // same reduction as the 2-leaf check: one barrier less on every fourth leaf
    float km[kK], kr[kK];
    s4_ld(s4_addr, km);
    s4_ld(s4_addr + kK, kr);
    tmem_wait_ld();
    if (act) {
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        const float diff = rho[j] - kr[j];
        s6[4] = fmaf(diff, km[j], s6[4]);
        s6[5] = fmaf(diff, m[j], s6[5]);
      }
    }
  }
  pf.mark(3);
  if (slot1) cx.reduce<6>(s6);
  else cx.reduce<4>(reinterpret_cast<float(&)[4]>(s6));
  pf.mark(4);
  bool ok = true;
  if (odd) {
    if (jmax >= 1) ok = (s6[2] >= 0.f) && (s6[3] >= 0.f);
    if (slot1 && jmax >= 2) ok = ok && (s6[4] >= 0.f) && (s6[5] >= 0.f);
    // the larger subtrees this leaf closes: slots [pc - ones, pc - 2), two checks per reduction
#pragma unroll 1
    for (int k = pc - ones; k < pc - 2; k += 2) {   // uniform trip count over the tile (shared leaf clock)
      const bool two = k + 1 < pc - 2;
      const int k1 = two ? k + 1 : k;
      float sd[4] = {0.f, 0.f, 0.f, 0.f};
      if (act)
        uturn_pair(e.ck_m + (size_t)k * kVS, e.ck_r + (size_t)k * kVS, e.ck_m + (size_t)k1 * kVS, e.ck_r + (size_t)k1 * kVS,
                   cl, rho, m, sd);
      cx.reduce<4>(sd);
      if (pc - k <= jmax) ok = ok && (sd[0] >= 0.f) && (sd[1] >= 0.f);
      if (two && pc - k1 <= jmax) ok = ok && (sd[2] >= 0.f) && (sd[3] >= 0.f);
    }
    // async kernel, last leaf of a 32-leaf chunk: the subtrees of 64, 128, .. leaves it closes start at the first
    // leaf of an earlier chunk of this lane's doubling
#pragma unroll 1
    for (int jj = 1; jj <= hi_checks; ++jj) {
      const bool chk = act && jj <= t_hi;
      float s2[2] = {0.f, 0.f};
      if (chk) {
        const int sl = __popc(ihi - (1 << jj) + 1);
        float vm[kK], vr[kK];
        seg_ldv(e.hi_m + (size_t)sl * kVS, cl, vm);
        seg_ldv(e.hi_r + (size_t)sl * kVS, cl, vr);
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const float diff = rho[j] - vr[j];
          s2[0] = fmaf(diff, vm[j], s2[0]);
          s2[1] = fmaf(diff, m[j], s2[1]);
        }
      }
      cx.reduce<2>(s2);
      if (chk) ok = ok && (s2[0] >= 0.f) && (s2[1] >= 0.f);
    }
  }
  pf.mark(5);
  if (act) {
    s.n += 1;
    s.slp = fmaf(0.5f, s6[0], e.lognorm);
    float en = s.slp - 0.5f * s6[1];                      // nuts.py:871-877
    en = isnan(en) ? -INFINITY : en;
    const float dH = en - H0;
    const bool nd_i = (-dH) < e.max_energy_diff;          // :880
    const float w_new = log_add_exp(s.bw, dH);            // :881-883
    if (lu_i <= (dH - w_new)) {                           // :897-901
      seg_stv(e.bx, cl, x);
      seg_stv(e.bg, cl, g);
      s.blp = s.slp; s.ben = en;
    }
    s.bw = w_new;
    if (nd_i) s.esum_sub += expf(fminf(dH, 0.f));         // :930-933 (act implies the chain continued so far)
    s.nd = s.nd && nd_i;                                  // :924-927,944
    s.alive = ok && nd_i;                                 // :921-922
    if (e.flags && s.alive) e.flags[gt & 3] = 1;
  }
  pf.mark(6);
  pf.leaf();
  return true;
}

// ---------------------------------------------------------------------------------------------
// tile_nuts_kernel: lock-step.  L2 scratch: other end, trajectory / subtree candidates, rho, checkpoint slots.
enum { kVOx = 0, kVOm, kVOg, kVCx, kVCg, kVBx, kVBg, kVRho, kVCk };   // checkpoints: kVCk + slot (m), + depth + slot (rho)

__global__ void __launch_bounds__(kThreads, 1)
tile_nuts_kernel(const ChainParams p, const DenseGaussianParams tp, float* __restrict__ scratch_all) {
  extern __shared__ __align__(128) unsigned char planes[];
  __shared__ Shared sh;
  __shared__ float lu[4][kM];   // log1p(-u) of the multinomial draws of 4 consecutive leaves
  Ctx cx;
  float* const dyn = reinterpret_cast<float*>(planes + 2 * kPlaneBytes);
  cx.init(&sh, planes, dyn + 2 * kVS, tp.P, tp.loc, tp.D, tp.scale);
  Prof pf;
  pf.init();
  const int D = tp.D;
  const int cl = cx.cl;
  const int nvec = kVCk + 2 * p.max_depth;
  float* const scr_s = scratch_all + (size_t)blockIdx.x * nvec * kVS + (size_t)(kK * cx.part) * kM;
  auto sv = [&](int v) -> float* { return scr_s + (size_t)v * kVS; };
  LeafEnv e;
  e.lc = sh.loc + kK * cx.part;
  e.bx = sv(kVBx); e.bg = sv(kVBg); e.ck_m = sv(kVCk); e.ck_r = sv(kVCk + p.max_depth);
  e.ckl = dyn + (size_t)(kK * cx.part) * kM;
  e.hi_m = nullptr; e.hi_r = nullptr;
  e.flags = sh.flags;
  e.unrolled = p.unrolled; e.lognorm = tp.lognorm; e.max_energy_diff = p.max_energy_diff;
  const int ntiles = (p.B + kM - 1) / kM;
  unsigned gt = 0;   // global leaf counter (rotates the "somebody continues" flags)
  for (int tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    const int c = tile_i * kM + cl;
    const bool live = c < p.B;
    const uint64_t cg = (uint64_t)p.chain_offset + (uint64_t)c;
    float x[kK], m[kK], g[kK], rho[kK];
    tile_load(p.x, c, D, cx.part, live, x);
    tile_load(p.g, c, D, cx.part, live, g);
    float lp = live ? p.lp[c] : 0.f;
    unsigned long long nleap_total = 0;
#pragma unroll 1
    for (int t = p.t0; t < p.t1; ++t) {
      const float eps_abs = p.step_kind == 0 ? p.step[0] : (live ? p.step[c] : 0.f);
      const uint32_t* sk = p.sched + (size_t)(t - p.t_sched0) * p.sched_stride;
      const uint32_t* hdr = sk + 2 * p.n_parts;
      const uint32_t* ku = hdr + 6 * p.max_depth;
      const int r = nuts_result_index(p, t);
      // ---- _start_trajectory_batched (nuts.py:512-539): momentum, H0; both ends, candidate, rho
      float s1[1] = {0.f};
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        const int d = kK * cx.part + j;
        const float mm = (live && d < D) ? nuts_momentum(p, sk, cg, d) : 0.f;
        m[j] = mm;
        s1[0] = fmaf(mm, mm, s1[0]);
      }
      seg_stv(sv(kVOx), cl, x); seg_stv(sv(kVOm), cl, m); seg_stv(sv(kVOg), cl, g);
      seg_stv(sv(kVCx), cl, x); seg_stv(sv(kVCg), cl, g);
      seg_stv(sv(kVRho), cl, m);
      cx.reduce<1>(s1);
      const float H0 = lp - 0.5f * s1[0];
      float slp = lp, olp = lp, clp = lp, cen = H0, cw = 0.f;
      float esum = 0.f;
      int nleap = 0;
      bool cont = live, notdiv = true, accepted = false, s_is_right = true;
      int any_cont = __syncthreads_or(cont ? 1 : 0);
#pragma unroll 1
      for (int it = 0; it < p.max_depth && any_cont; ++it) {
        // per-depth randoms of this chain (nuts.py:551-558, :622-625)
        Key kd{hdr[6 * it], hdr[6 * it + 1]}, kac{hdr[6 * it + 2], hdr[6 * it + 3]};
        const bool dir = (bits_at(kd, cg, (uint64_t)p.B_global, p.layout) & 1u) != 0;
        const float lacc = log1pf(-uniform_from_bits(bits_at(kac, cg, (uint64_t)p.B_global, p.layout), 0.f, 1.f));
        if (dir != s_is_right) {   // registers must hold the end that is extended
          float o[kK];
          seg_ldv(sv(kVOx), cl, o); seg_stv(sv(kVOx), cl, x);
#pragma unroll
          for (int j = 0; j < kK; ++j) x[j] = o[j];
          seg_ldv(sv(kVOm), cl, o); seg_stv(sv(kVOm), cl, m);
#pragma unroll
          for (int j = 0; j < kK; ++j) m[j] = o[j];
          seg_ldv(sv(kVOg), cl, o); seg_stv(sv(kVOg), cl, g);
#pragma unroll
          for (int j = 0; j < kK; ++j) g[j] = o[j];
          const float a = slp; slp = olp; olp = a;
          s_is_right = dir;
        }
        const float eps = dir ? eps_abs : -eps_abs;
        // _build_sub_tree init (nuts.py:713-791)
        seg_stv(e.bx, cl, x);
        seg_stv(e.bg, cl, g);
#pragma unroll
        for (int j = 0; j < kK; ++j) rho[j] = 0.f;
        LaneSub st;
        st.slp = slp; st.blp = slp; st.ben = slp; st.bw = -INFINITY; st.esum_sub = 0.f; st.n = 0;
        st.alive = cont; st.nd = notdiv;
        const int nsteps = 1 << it;
        const uint32_t* kud = ku + 2 * (nsteps - 1);
#pragma unroll 1
        for (int i = 0; i < nsteps; ++i, ++gt) {
          pf.mark(0);
          if (cx.half == 0 && (i & 3) == 0 && i + cx.slice < nsteps) {   // 4 leaves of multinomial uniforms
            Key kk{kud[2 * (i + cx.slice)], kud[2 * (i + cx.slice) + 1]};
            lu[cx.slice][cl] = log1pf(-uniform_from_bits(bits_at(kk, cg, (uint64_t)p.B_global, p.layout), 0.f, 1.f));
          }
          if (!nuts_leaf(cx, e, i, gt, st.alive, 31, -1, 0, 0, 0, eps, H0, lu[i & 3], x, m, g, rho, st, pf)) break;
        }
        slp = st.slp;
        const bool cont_f = st.alive;
        // _loop_tree_doubling tail (nuts.py:597-711)
        esum = st.esum_sub + esum;
        const float tw = cont_f ? st.bw : -INFINITY;
        const float wsum = log_add_exp(tw, cw);
        float thr = tw - cw;
        thr = isnan(thr) ? 0.f : thr;
        const bool swap = (lacc <= thr) && cont_f;
        cw = wsum;
        if (swap) {
          float o[kK];
          seg_ldv(e.bx, cl, o); seg_stv(sv(kVCx), cl, o);
          seg_ldv(e.bg, cl, o); seg_stv(sv(kVCg), cl, o);
          clp = st.blp; cen = st.ben;
        }
        float s2[2] = {0.f, 0.f};
        if (cont_f) {
          float rh[kK], om[kK];
          seg_ldv(sv(kVRho), cl, rh);
          seg_ldv(sv(kVOm), cl, om);
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            const float rr = rh[j] + rho[j];
            rh[j] = rr;
            s2[0] = fmaf(rr, m[j], s2[0]);
            s2[1] = fmaf(rr, om[j], s2[1]);
          }
          seg_stv(sv(kVRho), cl, rh);
        }
        cx.reduce<2>(s2);
        nleap += st.n;
        accepted = accepted || swap;
        notdiv = st.nd;
        cont = cont_f && (s2[0] >= 0.f) && (s2[1] >= 0.f);
        any_cont = __syncthreads_or(cont ? 1 : 0);       // nuts.py:404-407
      }
      // ---- results (nuts.py:424-445); the next state is the trajectory candidate
      seg_ldv(sv(kVCx), cl, x);
      seg_ldv(sv(kVCg), cl, g);
      lp = clp;
      const int leap = nleap * p.unrolled;
      nleap_total += (unsigned long long)leap;
      const float lar = logf(esum / (float)nleap);
      if (live && cx.part == 0 && p.lar_last) p.lar_last[c] = lar;
      if (r >= 0) {
        const Trace& tr = p.tr;
        if (tr.states) tile_store(tr.states, r, p.B, c, D, cx.part, live, x);
        if (tr.grads) tile_store(tr.grads, r, p.B, c, D, cx.part, live, g);
        if (live && cx.part == 0) {
          const size_t o = (size_t)r * p.B + c;
          if (tr.target_log_prob) tr.target_log_prob[o] = lp;
          if (tr.log_accept_ratio) tr.log_accept_ratio[o] = lar;
          if (tr.is_accepted) tr.is_accepted[o] = accepted ? 1 : 0;
          if (tr.leapfrogs_taken) tr.leapfrogs_taken[o] = leap;
          if (tr.has_divergence) tr.has_divergence[o] = notdiv ? 0 : 1;
          if (tr.reach_max_depth) tr.reach_max_depth[o] = cont ? 1 : 0;
          if (tr.energy) tr.energy[o] = cen;
          if (tr.step_size && c == 0 && p.step_kind == 0) tr.step_size[r] = p.step[0];
        }
      }
    }
    tile_store(p.x, 0, p.B, c, D, cx.part, live, x);
    tile_store(p.g, 0, p.B, c, D, cx.part, live, g);
    if (live && cx.part == 0) {
      p.lp[c] = lp;
      if (p.leapfrog_total) p.leapfrog_total[c] += nleap_total;
    }
  }
  cx.finish();
}

// ---------------------------------------------------------------------------------------------
// tile_nuts_async_kernel: every lane of the tile at its OWN position of its OWN tree (and its own transition) --
// no chain ever waits for another chain's deeper tree.
//
// The lock-step kernel makes all chains of a tile wait for the tile's deepest tree (measured utilisation 0.27
// on the 100-d ill-conditioned Gaussian: mean 276 of max 1023 leapfrogs).  Here the tile only shares a 32-tick
// CHUNK clock; tick i of a chunk is, for a lane in state
//   START : leaf i - 2^d of doubling d = floor(log2 i) of a new transition, i.e. doublings 0..4 sit in the
//           aligned blocks [1,2) [2,4) [4,8) [8,16) [16,32) of the chunk (tick 0 idles), and a doubling
//           ends / the next one begins after ticks 1, 3, 7, 15, 31;
//   CHUNK : leaf 32 q + i of a doubling d >= 5, q = the lane's own chunk counter.
// What must be uniform over a tile depends only on i: checkpoint write (even) vs U-turn check (odd), the popcount
// slot, how many of the 2-, 4-, .., 32-leaf subtrees close at this leaf (nuts.py:949-1071).  What differs per lane
// is masked: a START lane ignores closing subtrees larger than its doubling, a CHUNK lane adds the checks of its
// 64-, 128-, .. leaf subtrees at tick 31 against the checkpoints of its earlier chunks' first leaves.  All other
// work of a doubling / transition boundary (trajectory-level U-turn, candidate swap, direction draw, momentum
// draw, results) happens per lane at the chunk clock's boundary ticks.  A chain that U-turns inside a subtree
// idles until the end of its chunk (< 32 ticks).  A lane keeps its chain for all transitions of the launch and
// then takes the next unprocessed chain; every byte of a lane's state is private to the lane (registers, its
// columns of the CTA's scratch), so there are no records, queues or hand-offs.
constexpr int kS0 = 5;                       // doublings of the START state; chunks are 2^kS0 ticks
constexpr int kChunkTicks = 1 << kS0;
enum { kLaneNone = 0, kLaneStart = 1, kLaneChunk = 2 };
// CTA scratch vectors: other end, trajectory candidate, rho, subtree candidate, 5 local checkpoint slots (m, rho),
// max_depth - 5 per-lane slots of chunk-first leaves (m, rho)
enum { kAOx = 0, kAOm, kAOg, kACx, kACg, kARho, kABx, kABg, kACkM, kACkR = kACkM + kS0, kAHiM = kACkR + kS0 };

// FIFO of chains waiting for their next transition (ring buffer of B chain ids, -1 = empty slot): a lane hands its
// chain back after every transition and takes the chain at the head, so all chains advance at the same pace and
// every lane stays busy until the launch's last transitions (a chain's state between transitions is just x, g, lp
// in the chain-state arrays).
struct AsyncQueue {
  int* q;                      // [cap] ring of chain ids, -1 = empty slot
  unsigned long long* ctl;     // head (next ticket), tail (next push)
  int* t_next;                 // [B] next transition of each chain
  int cap;                     // >= B + lanes: live tickets (waiting lanes + queued chains) never share a slot
};
enum { kQHead = 0, kQTail = 1 };

__global__ void tile_async_init_kernel(AsyncQueue aq, int B, int t0) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) { aq.ctl[kQHead] = 0ull; aq.ctl[kQTail] = (unsigned long long)B; }
  if (c < aq.cap) aq.q[c] = c < B ? c : -1;
  if (c < B) aq.t_next[c] = t0;
}

static int async_scratch_vectors(int max_depth) {
  const int nhi = max_depth > kS0 ? max_depth - kS0 : 0;
  return kAHiM + 2 * nhi;
}

__global__ void __launch_bounds__(kThreads, 1)
tile_nuts_async_kernel(const ChainParams p, const DenseGaussianParams tp, float* __restrict__ scratch_all,
                       const AsyncQueue aq) {
  extern __shared__ __align__(128) unsigned char planes[];
  __shared__ Shared sh;
  __shared__ float lu[4][kM];
  __shared__ int new_chain[kM];
  __shared__ int hi_max, n_start;
  __shared__ int st_lane[kM], st_c[kM], st_t[kM];
  Ctx cx;
  float* const dyn = reinterpret_cast<float*>(planes + 2 * kPlaneBytes);
  cx.init(&sh, planes, dyn + 2 * kVS, tp.P, tp.loc, tp.D, tp.scale);
  Prof pf;
  pf.init();
  const int D = tp.D;
  const int cl = cx.cl;
  const int nhi = p.max_depth > kS0 ? p.max_depth - kS0 : 0;
  const int nvec = kAHiM + 2 * nhi;
  float* const scr_s = scratch_all + (size_t)blockIdx.x * nvec * kVS + (size_t)(kK * cx.part) * kM;
  auto sv = [&](int v) -> float* { return scr_s + (size_t)v * kVS; };
  LeafEnv e;
  e.lc = sh.loc + kK * cx.part;
  e.bx = sv(kABx); e.bg = sv(kABg); e.ck_m = sv(kACkM); e.ck_r = sv(kACkR);
  e.ckl = dyn + (size_t)(kK * cx.part) * kM;
  e.hi_m = sv(kAHiM); e.hi_r = sv(kAHiM + nhi);
  e.flags = nullptr;
  e.unrolled = p.unrolled; e.lognorm = tp.lognorm; e.max_energy_diff = p.max_energy_diff;
  const uint64_t Bg = (uint64_t)p.B_global;

  // ---- lane state
  int c = -1;
  int t = p.t0;
  int type = kLaneNone;
  bool fin = false;            // the transition is complete; results are emitted at the end of the chunk
  int it = 0, ihi = 0;
  float lp = 0.f, H0 = 0.f, olp = 0.f, clp = 0.f, cen = 0.f, cw = 0.f, esum = 0.f;
  int nleap = 0;
  bool cont = false, notdiv = true, accepted = false, s_is_right = true;
  LaneSub st;
  st.slp = st.blp = st.ben = st.bw = st.esum_sub = 0.f; st.n = 0; st.alive = false; st.nd = true;
  float eps_abs = 0.f, eps = 0.f;
  float x[kK], m[kK], g[kK], rho[kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) { x[j] = 0.f; m[j] = 0.f; g[j] = 0.f; rho[j] = 0.f; }
  const unsigned long long qcap = (unsigned long long)aq.cap;
  const long long total_pushes = (long long)p.B * (long long)(p.t1 - p.t0);
  long long ticket = -1;   // (part 0 of a lane) my position in the FIFO while the lane waits for a chain

  // per-transition key schedule of this lane
  auto keys = [&](const uint32_t*& sk, const uint32_t*& hdr, const uint32_t*& ku) {
    sk = p.sched + (size_t)(t - p.t_sched0) * p.sched_stride;
    hdr = sk + 2 * p.n_parts;
    ku = hdr + 6 * p.max_depth;
  };

  // ---- begin doubling `it` of this lane (nuts.py:551-558 direction, :713-791 _build_sub_tree init)
  auto begin_doubling = [&]() {
    const uint32_t *sk, *hdr, *ku;
    keys(sk, hdr, ku);
    Key kd{hdr[6 * it], hdr[6 * it + 1]};
    const bool dir = (bits_at(kd, (uint64_t)p.chain_offset + (uint64_t)c, Bg, p.layout) & 1u) != 0;
    if (dir != s_is_right) {   // registers must hold the end that is extended
      float o[kK];
      seg_ldv(sv(kAOx), cl, o); seg_stv(sv(kAOx), cl, x);
#pragma unroll
      for (int j = 0; j < kK; ++j) x[j] = o[j];
      seg_ldv(sv(kAOm), cl, o); seg_stv(sv(kAOm), cl, m);
#pragma unroll
      for (int j = 0; j < kK; ++j) m[j] = o[j];
      seg_ldv(sv(kAOg), cl, o); seg_stv(sv(kAOg), cl, g);
#pragma unroll
      for (int j = 0; j < kK; ++j) g[j] = o[j];
      const float a = st.slp; st.slp = olp; olp = a;
      s_is_right = dir;
    }
    seg_stv(e.bx, cl, x);
    seg_stv(e.bg, cl, g);
#pragma unroll
    for (int j = 0; j < kK; ++j) rho[j] = 0.f;
    st.blp = st.slp; st.ben = st.slp; st.bw = -INFINITY; st.esum_sub = 0.f; st.n = 0; st.nd = notdiv; st.alive = true;
    ihi = 0;
    eps = dir ? eps_abs : -eps_abs;
  };

  // ---- start transition t of chain c from (x, g, lp): _start_trajectory_batched (nuts.py:512-539).  Tile-uniform
  // (one cross-part reduction); `go` selects the lanes that start.
  auto start_transitions = [&](bool go) {
    // momentum ~ N(0, I).  After the first chunk only a few lanes start at a time: their draws are spread over the
    // whole CTA and handed over through shared memory (the previous-leaf checkpoint buffer is dead between chunks)
    if (threadIdx.x == 0) n_start = 0;
    __syncthreads();
    if (go && cx.part == 0) {
      const int idx = atomicAdd(&n_start, 1);
      st_lane[idx] = cl; st_c[idx] = c; st_t[idx] = t;
    }
    __syncthreads();
    float* const mom = dyn;   // [kKP][kM]
    for (int w = threadIdx.x; w < n_start * D; w += kThreads) {
      const int li = w / D, d = w - li * D;
      const uint32_t* skl = p.sched + (size_t)(st_t[li] - p.t_sched0) * p.sched_stride;
      mom[d * kM + st_lane[li]] = nuts_momentum(p, skl, (uint64_t)p.chain_offset + (uint64_t)st_c[li], d);
    }
    __syncthreads();
    pf.mark(12);
    float s1[1] = {0.f};
    if (go) {
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        const int d = kK * cx.part + j;
        const float mm = d < D ? mom[d * kM + cl] : 0.f;
        m[j] = mm;
        s1[0] = fmaf(mm, mm, s1[0]);
      }
      seg_stv(sv(kAOx), cl, x); seg_stv(sv(kAOm), cl, m); seg_stv(sv(kAOg), cl, g);
      seg_stv(sv(kACx), cl, x); seg_stv(sv(kACg), cl, g);
      seg_stv(sv(kARho), cl, m);
    }
    cx.reduce<1>(s1);
    if (go) {
      H0 = lp - 0.5f * s1[0];
      st.slp = lp; olp = lp; clp = lp; cen = H0; cw = 0.f; esum = 0.f;
      nleap = 0;
      cont = true; notdiv = true; accepted = false; s_is_right = true;
      type = kLaneStart; fin = false; it = 0;
      begin_doubling();
    }
  };

  // ---- end the current doubling of the selected lanes (_loop_tree_doubling tail, nuts.py:597-711) and begin the
  // next one, or mark the transition finished.  Tile-uniform (one cross-part reduction).
  auto doubling_boundary = [&](bool endd, bool chunk_end) {
    float s2[2] = {0.f, 0.f};
    bool swap = false;
    if (endd) {
      const uint32_t *sk, *hdr, *ku;
      keys(sk, hdr, ku);
      Key kac{hdr[6 * it + 2], hdr[6 * it + 3]};
      const float lacc =
          log1pf(-uniform_from_bits(bits_at(kac, (uint64_t)p.chain_offset + (uint64_t)c, Bg, p.layout), 0.f, 1.f));
      esum = st.esum_sub + esum;
      const float tw = st.alive ? st.bw : -INFINITY;
      const float wsum = log_add_exp(tw, cw);
      float thr = tw - cw;
      thr = isnan(thr) ? 0.f : thr;
      swap = (lacc <= thr) && st.alive;
      cw = wsum;
      if (swap) {
        float o[kK];
        seg_ldv(e.bx, cl, o); seg_stv(sv(kACx), cl, o);
        seg_ldv(e.bg, cl, o); seg_stv(sv(kACg), cl, o);
        clp = st.blp; cen = st.ben;
      }
      if (st.alive) {
        float rh[kK], om[kK];
        seg_ldv(sv(kARho), cl, rh);
        seg_ldv(sv(kAOm), cl, om);
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const float rr = rh[j] + rho[j];
          rh[j] = rr;
          s2[0] = fmaf(rr, m[j], s2[0]);
          s2[1] = fmaf(rr, om[j], s2[1]);
        }
        seg_stv(sv(kARho), cl, rh);
      }
    }
    cx.reduce<2>(s2);
    if (endd) {
      nleap += st.n;
      accepted = accepted || swap;
      notdiv = st.nd;
      cont = st.alive && (s2[0] >= 0.f) && (s2[1] >= 0.f);
      if (cont && it + 1 < p.max_depth) {
        it += 1;
        if (chunk_end && type == kLaneStart) type = kLaneChunk;
        begin_doubling();
      } else {
        fin = true;
        st.alive = false;
      }
    }
  };

  while (true) {
    // ------------------------------------------------------------ idle lanes take the chains at the head of the FIFO
    // A lane without a chain draws a TICKET (its position in the FIFO, one atomicAdd) and looks at that slot once
    // per chunk: no lane ever spins and nobody needs the tail.  Tickets beyond the total number of pushes
    // (every chain is pushed once per transition) will never be served: the lane retires.
    {
      if (threadIdx.x == 0) hi_max = 0;
      const bool want = type == kLaneNone;
      if (want && cx.part == 0) {
        if (ticket < 0) ticket = (long long)atomicAdd(aq.ctl + kQHead, 1ull);
        int nc = -1;
        if (ticket < total_pushes) {
          volatile int* slot = aq.q + (size_t)((unsigned long long)ticket % qcap);
          nc = *slot;
          if (nc >= 0) { *slot = -1; ticket = -1; }
        }
        new_chain[cl] = nc;
      }
      __syncthreads();   // (the chain's state is read with ld.cg from L2, where its last owner's fenced stores are)
      bool go = false;
      if (want) {
        c = new_chain[cl];
        if (c >= 0) {
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            const bool in = kK * cx.part + j < D;
            x[j] = in ? __ldcg(p.x + (size_t)c * D + kK * cx.part + j) : 0.f;
            g[j] = in ? __ldcg(p.g + (size_t)c * D + kK * cx.part + j) : 0.f;
          }
          lp = __ldcg(p.lp + c);
          eps_abs = p.step_kind == 0 ? p.step[0] : p.step[c];
          t = __ldcg(aq.t_next + c);
          go = true;
        }
      }
      pf.mark(11);
      start_transitions(go);
    }
    pf.mark(10);
    // ------------------------------------------------------------ chunk set-up
    const int has_start = __syncthreads_or(type == kLaneStart ? 1 : 0);
    const int any_work = __syncthreads_or(type != kLaneNone ? 1 : 0);
    if (!any_work) {
      // lanes that still hold a servable ticket wait for their chain; otherwise every lane of the tile has retired
      const int pending = __syncthreads_or(cx.part == 0 && ticket >= 0 && ticket < total_pushes ? 1 : 0);
      if (!pending) break;
      __nanosleep(2000);
      continue;
    }
    const uint32_t *sk, *hdr, *ku;
    keys(sk, hdr, ku);
    const uint64_t cg = (uint64_t)p.chain_offset + (uint64_t)(c >= 0 ? c : 0);
    // multinomial key of tick i: kb[2 i] (START: leaf i - 2^d of doubling d has key (2^d - 1) + i - 2^d = i - 1)
    const uint32_t* kb = type == kLaneChunk ? ku + 2 * ((1 << it) - 1) + 2 * (ihi * kChunkTicks) : ku - 2;
    const int nchunks = type == kLaneChunk ? 1 << (it - kS0) : 1;
    const int t_hi = type == kLaneChunk ? __ffs(~ihi) - 1 : 0;
    // leaf 0 of this chunk is checked again by later chunks of the doubling -> keep it in the lane's slot popc(ihi)
    const int hi_slot_w = (type == kLaneChunk && (ihi & 1) == 0 && ihi + 1 < nchunks) ? __popc(ihi) : -1;
    if (cx.part == 0 && t_hi > 0) atomicMax(&hi_max, t_hi);
    __syncthreads();
    const int hi_checks = hi_max;

#pragma unroll 1
    for (int i = 0; i < kChunkTicks; ++i) {
      pf.mark(0);
      const bool act = st.alive && !fin && (type == kLaneChunk || (type == kLaneStart && i >= 1));
      if (cx.half == 0 && (i & 3) == 0) {   // 4 ticks of multinomial uniforms, one per slice
        const int ii = i + cx.slice;
        if (type == kLaneChunk || (type == kLaneStart && ii >= 1)) {
          Key kk{kb[2 * ii], kb[2 * ii + 1]};
          lu[cx.slice][cl] = log1pf(-uniform_from_bits(bits_at(kk, cg, Bg, p.layout), 0.f, 1.f));
        }
      }
      // the largest subtree (2^jmax leaves) that can close inside this lane's doubling at this tick
      const int jmax = type == kLaneChunk ? kS0 : 31 - __clz(i | 1);
      nuts_leaf(cx, e, i, 0u, act, jmax, hi_slot_w, ihi, t_hi, i == kChunkTicks - 1 ? hi_checks : 0, eps, H0, lu[i & 3],
                x, m, g, rho, st, pf);
      // START lanes: doublings 0..3 end after ticks 1, 3, 7, 15
      if (has_start && i >= 1 && i < kChunkTicks - 1 && ((i + 1) & i) == 0) {
        pf.mark(0);
        doubling_boundary(type == kLaneStart && !fin, false);
        pf.mark(7);
      }
    }
    pf.mark(0);
    // ------------------------------------------------------------ end of the chunk
    {
      const bool endd = !fin && (type == kLaneStart || (type == kLaneChunk && (!st.alive || ihi + 1 == nchunks)));
      if (type == kLaneChunk && !endd) ihi += 1;
      doubling_boundary(endd, true);
    }
    pf.mark(8);
    // finished transitions: results (nuts.py:424-445; the next state is the trajectory candidate); the chain goes
    // back to the FIFO
    {
      const bool done = fin && type != kLaneNone;
      if (done) {
        seg_ldv(sv(kACx), cl, x);
        seg_ldv(sv(kACg), cl, g);
        lp = clp;
        const int leap = nleap * p.unrolled;
        const float lar = logf(esum / (float)nleap);
        if (cx.part == 0 && p.lar_last) p.lar_last[c] = lar;
        const int r = nuts_result_index(p, t);
        if (r >= 0) {
          const Trace& tr = p.tr;
          if (tr.states) tile_store(tr.states, r, p.B, c, D, cx.part, true, x);
          if (tr.grads) tile_store(tr.grads, r, p.B, c, D, cx.part, true, g);
          if (cx.part == 0) {
            const size_t o = (size_t)r * p.B + c;
            if (tr.target_log_prob) tr.target_log_prob[o] = lp;
            if (tr.log_accept_ratio) tr.log_accept_ratio[o] = lar;
            if (tr.is_accepted) tr.is_accepted[o] = accepted ? 1 : 0;
            if (tr.leapfrogs_taken) tr.leapfrogs_taken[o] = leap;
            if (tr.has_divergence) tr.has_divergence[o] = notdiv ? 0 : 1;
            if (tr.reach_max_depth) tr.reach_max_depth[o] = cont ? 1 : 0;
            if (tr.energy) tr.energy[o] = cen;
            if (tr.step_size && c == 0 && p.step_kind == 0) tr.step_size[r] = p.step[0];
          }
        }
        t += 1;
        tile_store(p.x, 0, p.B, c, D, cx.part, true, x);
        tile_store(p.g, 0, p.B, c, D, cx.part, true, g);
        if (cx.part == 0) {
          p.lp[c] = lp;
          aq.t_next[c] = t;
          if (p.leapfrog_total) atomicAdd(p.leapfrog_total + c, (unsigned long long)leap);
        }
        type = kLaneNone; fin = false; st.alive = false;
      }
      __threadfence();
      __syncthreads();   // all eight parts of every finished chain have written its state
      if (done && cx.part == 0 && t < p.t1) {
        const unsigned long long sl = atomicAdd(aq.ctl + kQTail, 1ull);
        *(volatile int*)(aq.q + (size_t)(sl % qcap)) = c;
      }
    }
    pf.mark(9);
  }
  cx.finish();
}

#ifdef PB2_TILE_PROF
static void dump_tile_prof(pb2_ctx* ctx) {
  unsigned long long h[2][16];
  cudaStreamSynchronize(ctx->stream);
  cudaMemcpyFromSymbol(h, g_tile_prof, sizeof(h));
  static const char* nm[13] = {"head", "kick+stage", "contract", "post", "reduce4", "extra checks", "scalars+take",
                               "START boundaries", "chunk-end boundary", "finish", "start(rest)", "acquire", "momentum"};
  for (int w = 0; w < 2; ++w) {
    fprintf(stderr, "[tileprof t%d] leaves %llu:", w ? 511 : 0, h[w][15]);
    for (int k = 0; k < 13; ++k) fprintf(stderr, " %s %.0f", nm[k], h[w][15] ? (double)h[w][k] / h[w][15] : 0.0);
    fprintf(stderr, "\n");
  }
  unsigned long long z[2][16] = {};
  cudaMemcpyToSymbol(g_tile_prof, z, sizeof(z));
}
#else
static void dump_tile_prof(pb2_ctx*) {}
#endif

static int ensure_scratch(pb2_ctx* ctx, size_t need, const char* what) {
  if (need <= ctx->ckpt_bytes) return PB2_OK;
  if (ctx->d_ckpt) cudaFree(ctx->d_ckpt);
  ctx->d_ckpt = nullptr;
  ctx->ckpt_bytes = 0;
  if (int rc = check_cuda(ctx, cudaMalloc(&ctx->d_ckpt, need), what)) return rc;
  ctx->ckpt_bytes = need;
  return PB2_OK;
}

int launch_tile_nuts(pb2_ctx* ctx, const pb2_target* tgt, ChainParams& p) {
  DenseGaussianParams tp{tgt->d_a, tgt->d_b, tgt->scalar, tgt->dim};
  tp.scale = p.scale;
  // P hi/lo planes + the previous leaf's checkpoint (momentum, rho) + the gradient exchange buffer
  const size_t smem = 2 * (size_t)kPlaneBytes + (2 * kVS + kXbufFloats) * sizeof(float);
  const int ntiles = (p.B + kM - 1) / kM;
  // every lane at its own position of its own tree (the lock-step kernel remains as the literal batched algorithm:
  // dense_variant 3, or max_tree_depth <= 5)
  if (ctx->dense_variant != 3 && p.max_depth > kS0) {
    const int agrid = std::min(ntiles, getenv("PB2_ASYNC_GRID") ? atoi(getenv("PB2_ASYNC_GRID")) : ctx->num_sms);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t scr_bytes = up((size_t)agrid * async_scratch_vectors(p.max_depth) * kVS * sizeof(float));
    AsyncQueue aq;
    aq.cap = p.B + agrid * kM + 64;
    const size_t q_bytes = up((size_t)aq.cap * sizeof(int));
    if (int rc = ensure_scratch(ctx, scr_bytes + 2 * q_bytes + 256, "cudaMalloc(tile async scratch)")) return rc;
    unsigned char* base = reinterpret_cast<unsigned char*>(ctx->d_ckpt);
    aq.q = reinterpret_cast<int*>(base + scr_bytes);
    aq.t_next = reinterpret_cast<int*>(base + scr_bytes + q_bytes);
    aq.ctl = reinterpret_cast<unsigned long long*>(base + scr_bytes + 2 * q_bytes);
    tile_async_init_kernel<<<(aq.cap + 255) / 256, 256, 0, ctx->stream>>>(aq, p.B, p.t0);
    if (int rc = check_cuda(ctx, cudaFuncSetAttribute(tile_nuts_async_kernel,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(tile_nuts_async)"))
      return rc;
    tile_nuts_async_kernel<<<agrid, kThreads, smem, ctx->stream>>>(p, tp, ctx->d_ckpt, aq);
    ctx->launches += 2;
    dump_tile_prof(ctx);
    return check_cuda(ctx, cudaGetLastError(), "tile_nuts_async_kernel");
  }
  const int grid = std::min(ntiles, ctx->num_sms);
  const size_t need = (size_t)(kVCk + 2 * p.max_depth) * kVS * sizeof(float) * grid;
  if (int rc = ensure_scratch(ctx, need, "cudaMalloc(tile scratch)")) return rc;
  if (int rc = check_cuda(ctx, cudaFuncSetAttribute(tile_nuts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)smem), "cudaFuncSetAttribute(tile_nuts)"))
    return rc;
  tile_nuts_kernel<<<grid, kThreads, smem, ctx->stream>>>(p, tp, ctx->d_ckpt);
  ctx->launches += 1;
  dump_tile_prof(ctx);
  return check_cuda(ctx, cudaGetLastError(), "tile_nuts_kernel");
}

}  // namespace pb2
