// 128-chain TILE NUTS kernels for the dense-Gaussian target (tcgen05 path, pb2_tile128.cuh): a chain = one MMA row =
// two threads with x, m, rho in registers and the gradient in the TMEM accumulator.
//   tile128_nuts_kernel       : NoUTurnSampler.one_step (tfp/mcmc/nuts.py:321-946) for a tile run in LOCK-STEP, i.e.
//                               literally the reference's batched algorithm (shared doubling / leaf counters,
//                               per-chain masks) -- max_tree_depth <= 5, or dense_variant 3 (the bit-exact partner of
//                               the asynchronous kernel in the tests);
//   tile128_nuts_async_kernel : every lane at its own position of its own tree and transition -- all other launches
//                               (single transitions, adaptation steps and fused multi-transition runs).
// Both call the same nuts_leaf(), so a chain's arithmetic is the same instruction sequence in both and the
// results are bit-identical (tests/test_gpu_parity.py).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "pb2_tile128.cuh"

namespace pb2 {
namespace t128 {
using namespace tile128;

#ifdef PB2_TILE_PROF
__device__ unsigned long long g_tile_prof[2][16];
struct Prof {
  int w;
  long long t;
  __device__ void init() {
    w = (blockIdx.x == 0 && threadIdx.x == 0) ? 0 : ((blockIdx.x == 0 && threadIdx.x == kWorkers - 1) ? 1 : -1);
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
// Where my 52 dims live: registers x, m (moving end), rho (cumulative momentum of the subtree); TMEM D columns =
// gradient at the moving end; shared memory ckl = the checkpoint written by the previous (even) leaf; L2 scratch = the
// popcount-indexed checkpoint slots that later leaves need (only leaves with i % 4 == 0 are read again after leaf
// i + 1) and the subtree candidate.
struct LaneSub {
  float slp;        // log-prob at the moving end
  float blp, ben;   // subtree candidate's log-prob and energy
  float bw;         // log-sum of the subtree's weights
  float esum_sub;   // sum of min(1, exp(dH)) over the subtree's leaves (for log_accept_ratio)
  int n;            // leaves taken
  bool alive;       // no U-turn / divergence inside the subtree so far
  bool nd;          // not diverged
};

// all vector pointers = the vector's base + my first part block
struct LeafEnv {
  const float* lc;        // loc of my half (shared memory)
  float* bx;              // subtree candidate (x, g)
  float* bg;
  float* ck_m;            // checkpoint slot k: momentum at ck_m + k * kVS, rho at ck_r + k * kVS
  float* ck_r;
  float* ckl;             // shared memory: previous even leaf's checkpoint (momentum; rho at + kVS)
  float* hi_m;            // per-lane slots of chunk-first leaves (async kernel)
  float* hi_r;
  int* flags;             // lock-step kernel: "some chain continues" flags (nullptr in the async kernel)
  int unrolled;
  float lognorm, max_energy_diff;
  uint64_t Bg;
  int layout;
};

// my four parts' sum in the order of the pair tree
__device__ __forceinline__ float sum4(const float (&p)[kPT]) { return (p[0] + p[1]) + (p[2] + p[3]); }

// returns false if the lock-step tile stopped (no chain continues)
template <bool kLock>
__device__ __forceinline__ bool nuts_leaf(Ctx& cx, const LeafEnv& e, int i, unsigned gt, bool act, int jmax, int hi_slot_w,
                                          int ihi, int t_hi, int hi_checks, float eps, float H0, const uint32_t* kk,
                                          uint64_t cg, float (&x)[kDT], float (&m)[kDT], float (&rho)[kDT], LaneSub& s,
                                          Prof& pf) {
  const int cl = cx.cl;
  const float heps = 0.5f * eps;
  float lu_i = 0.f;
  // one leapfrog (leapfrog_integrator.py:280-309 with L = unrolled_leapfrog_steps)
#pragma unroll 1
  for (int l = 0; l < e.unrolled; ++l) {
    for_parts([&](auto pp) {
      constexpr int PP = decltype(pp)::value;
      if (l == 0) {   // first half kick with the gradient the previous leaf left in D
        float g[kK];
        cx.ld_g<PP>(g);
        if (act) {
#pragma unroll
          for (int j = 0; j < kK; ++j) m[kK * PP + j] = m[kK * PP + j] + heps * g[j];
        }
      }
      if (act) {
#pragma unroll
        for (int j = 0; j < kK; ++j) x[kK * PP + j] = x[kK * PP + j] + eps * m[kK * PP + j];
      }
      cx.stage_part<PP>(x + kK * PP);
    });
    pf.mark(1);
    cx.contract_begin();
    if (l == 0) {
      // in the contraction's shadow: the multinomial uniform of this leaf (nuts.py:897-901)
      if (kk) {
        Key key{kk[0], kk[1]};
        lu_i = log1pf(-uniform_from_bits(bits_at(key, cg, e.Bg, e.layout), 0.f, 1.f));
      }
    }
    cx.contract_end();
    pf.mark(2);
    if (kLock && l == 0) {
      // flag raised at the end of the previous leaf; leaving mid-leaf is harmless because no chain of the
      // tile continues (ends / candidates are final)
      Ctx::wsync();
      if (i > 0 && e.flags[(gt - 1) & 3] == 0) return false;
      if (threadIdx.x == 0) e.flags[(gt + 1) & 3] = 0;
    }
    if (l + 1 < e.unrolled) {
      for_parts([&](auto pp) {
        constexpr int PP = decltype(pp)::value;
        float g[kK];
        cx.ld_g<PP>(g);
        if (act) {
#pragma unroll
          for (int j = 0; j < kK; ++j) m[kK * PP + j] = m[kK * PP + j] + eps * g[j];
        }
      });
    }
  }
  // full kick of the last leapfrog, last half kick back, rho_subtree, checkpoint store / U-turn checks of the closing
  // 2- and 4-leaf subtrees (nuts.py:826-869, 949-1010)
  // partials: <x - mu, g>, |m|^2, dots vs the previous leaf, dots vs slot pc - 2
  float q0[kPT], q1[kPT], q2[kPT], q3[kPT], q4[kPT], q5[kPT];
  const int pc = __popc(i);
  const bool odd = (i & 1) != 0;
  const int ones = __ffs(~i) - 1;          // trailing ones: the leaf closes subtrees of 2, 4, .., 2^ones leaves
  const bool slot1 = odd && ones >= 2;     // (tile-uniform) the 4-leaf subtree closes: its checkpoint is slot pc - 2
  for_parts([&](auto pp) {
    constexpr int PP = decltype(pp)::value;
    float g[kK];
    cx.ld_g<PP>(g);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
    if (act) {
      float* mp = m + kK * PP;
      float* rp = rho + kK * PP;
      const float* xp = x + kK * PP;
      const float* lc = e.lc + kK * PP;
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        mp[j] = mp[j] + eps * g[j];
        mp[j] = mp[j] - heps * g[j];
        a0 = fmaf(xp[j] - lc[j], g[j], a0);
        a1 = fmaf(mp[j], mp[j], a1);
      }
      if (!odd) {
        seg_stp(e.ckl + PP * kPartBlk, cl, mp);
        seg_stp(e.ckl + kVS + PP * kPartBlk, cl, rp);
        if ((i & 3) == 0) {   // an even leaf that is checked again after leaf i + 1
          seg_stp(e.ck_m + (size_t)pc * kVS + PP * kPartBlk, cl, mp);
          seg_stp(e.ck_r + (size_t)pc * kVS + PP * kPartBlk, cl, rp);
        }
        if (i == 0 && hi_slot_w >= 0) {
          seg_stp(e.hi_m + (size_t)hi_slot_w * kVS + PP * kPartBlk, cl, mp);
          seg_stp(e.hi_r + (size_t)hi_slot_w * kVS + PP * kPartBlk, cl, rp);
        }
      }
#pragma unroll
      for (int j = 0; j < kK; ++j) rp[j] = rp[j] + mp[j];
      if (odd) {
        float km[kK], kr[kK];
        seg_ldp(e.ckl + PP * kPartBlk, cl, km);
        seg_ldp(e.ckl + kVS + PP * kPartBlk, cl, kr);
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const float diff = rp[j] - kr[j];
          a2 = fmaf(diff, km[j], a2);
          a3 = fmaf(diff, mp[j], a3);
        }
        if (slot1) {   // same reduction: one exchange less on every fourth leaf
          seg_ldp(e.ck_m + (size_t)(pc - 2) * kVS + PP * kPartBlk, cl, km);
          seg_ldp(e.ck_r + (size_t)(pc - 2) * kVS + PP * kPartBlk, cl, kr);
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            const float diff = rp[j] - kr[j];
            a4 = fmaf(diff, km[j], a4);
            a5 = fmaf(diff, mp[j], a5);
          }
        }
      }
    }
    q0[PP] = a0; q1[PP] = a1; q2[PP] = a2; q3[PP] = a3; q4[PP] = a4; q5[PP] = a5;
  });
  float s6[6] = {sum4(q0), sum4(q1), sum4(q2), sum4(q3), sum4(q4), sum4(q5)};
  pf.mark(3);
  if (slot1) cx.reduce<6>(s6);
  else cx.reduce<4>(reinterpret_cast<float(&)[4]>(s6));
  pf.mark(4);
  bool ok = true;
  if (odd) {
    if (jmax >= 1) ok = (s6[2] >= 0.f) && (s6[3] >= 0.f);
    if (slot1 && jmax >= 2) ok = ok && (s6[4] >= 0.f) && (s6[5] >= 0.f);
    // the larger subtrees this leaf closes: slots [pc - ones, pc - 2), two checks per exchange
    // (nuts.py:949-1010: s[2q] = <rho - rho_k, m_k>, s[2q+1] = <rho - rho_k, m>)
#pragma unroll 1
    for (int k = pc - ones; k < pc - 2; k += 2) {   // uniform trip count over the tile (shared leaf clock)
      const bool two = k + 1 < pc - 2;
      const int k1 = two ? k + 1 : k;
      float d0[kPT], d1[kPT], d2[kPT], d3[kPT];
      for_parts([&](auto pp) {
        constexpr int PP = decltype(pp)::value;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (act) {
          const float* rp = rho + kK * PP;
          const float* mp = m + kK * PP;
          float ra[kK], ma[kK], rb[kK], mb[kK];
          seg_ldp(e.ck_r + (size_t)k * kVS + PP * kPartBlk, cl, ra);
          seg_ldp(e.ck_m + (size_t)k * kVS + PP * kPartBlk, cl, ma);
          seg_ldp(e.ck_r + (size_t)k1 * kVS + PP * kPartBlk, cl, rb);
          seg_ldp(e.ck_m + (size_t)k1 * kVS + PP * kPartBlk, cl, mb);
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            const float e0 = rp[j] - ra[j], e1 = rp[j] - rb[j];
            a0 = fmaf(e0, ma[j], a0);
            a1 = fmaf(e0, mp[j], a1);
            a2 = fmaf(e1, mb[j], a2);
            a3 = fmaf(e1, mp[j], a3);
          }
        }
        d0[PP] = a0; d1[PP] = a1; d2[PP] = a2; d3[PP] = a3;
      });
      float sd[4] = {sum4(d0), sum4(d1), sum4(d2), sum4(d3)};
      cx.reduce<4>(sd);
      if (pc - k <= jmax) ok = ok && (sd[0] >= 0.f) && (sd[1] >= 0.f);
      if (two && pc - k1 <= jmax) ok = ok && (sd[2] >= 0.f) && (sd[3] >= 0.f);
    }
    // async kernel, last leaf of a 32-leaf chunk: the subtrees of 64, 128, .. leaves it closes start at the first
    // leaf of an earlier chunk of this lane's doubling
#pragma unroll 1
    for (int jj = 1; jj <= hi_checks; ++jj) {
      const bool chk = act && jj <= t_hi;
      float d0[kPT], d1[kPT];
      const int sl = chk ? __popc(ihi - (1 << jj) + 1) : 0;
      for_parts([&](auto pp) {
        constexpr int PP = decltype(pp)::value;
        float a0 = 0.f, a1 = 0.f;
        if (chk) {
          const float* rp = rho + kK * PP;
          const float* mp = m + kK * PP;
          float vm[kK], vr[kK];
          seg_ldp(e.hi_m + (size_t)sl * kVS + PP * kPartBlk, cl, vm);
          seg_ldp(e.hi_r + (size_t)sl * kVS + PP * kPartBlk, cl, vr);
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            const float diff = rp[j] - vr[j];
            a0 = fmaf(diff, vm[j], a0);
            a1 = fmaf(diff, mp[j], a1);
          }
        }
        d0[PP] = a0; d1[PP] = a1;
      });
      float s2[2] = {sum4(d0), sum4(d1)};
      cx.reduce<2>(s2);
      if (chk) ok = ok && (s2[0] >= 0.f) && (s2[1] >= 0.f);
    }
  }
  pf.mark(5);
  bool take = false;
  if (act) {
    s.n += 1;
    s.slp = fmaf(0.5f, s6[0], e.lognorm);
    float en = s.slp - 0.5f * s6[1];                      // nuts.py:871-877
    en = isnan(en) ? -INFINITY : en;
    const float dH = en - H0;
    const bool nd_i = (-dH) < e.max_energy_diff;          // :880
    const float w_new = log_add_exp(s.bw, dH);            // :881-883
    if (lu_i <= (dH - w_new)) {                           // :897-901
      take = true;
      s.blp = s.slp; s.ben = en;
    }
    s.bw = w_new;
    if (nd_i) s.esum_sub += expf(fminf(dH, 0.f));         // :930-933 (act implies the chain continued so far)
    s.nd = s.nd && nd_i;                                  // :924-927,944
    s.alive = ok && nd_i;                                 // :921-922
    if (kLock && s.alive) e.flags[gt & 3] = 1;
  }
  if (__any_sync(kFull, take)) {   // the leaf becomes the subtree's candidate (TMEM loads are warp-collective)
    for_parts([&](auto pp) {
      constexpr int PP = decltype(pp)::value;
      float g[kK];
      cx.ld_g<PP>(g);
      if (take) {
        seg_stp(e.bx + PP * kPartBlk, cl, x + kK * PP);
        seg_stp(e.bg + PP * kPartBlk, cl, g);
      }
    });
  }
  pf.mark(6);
  pf.leaf();
  return true;
}

// Exchange the moving end (registers x, m; D columns g) with the other end (scratch ox, om, og) for the lanes with
// `sw`, and start the subtree of the lanes with `doit`: subtree candidate = the moving end, rho = 0.  Warp-uniform
// call (TMEM accesses are warp-collective).
__device__ __forceinline__ void swap_and_seed(const Ctx& cx, bool doit, bool sw, float* ox, float* om, float* og, float* bx,
                                              float* bg, float (&x)[kDT], float (&m)[kDT], float (&rho)[kDT]) {
  const int cl = cx.cl;
  const bool any_sw = __any_sync(kFull, sw);
  if (!__any_sync(kFull, doit)) return;
  for_parts([&](auto pp) {
    constexpr int PP = decltype(pp)::value;
    float g[kK];
    cx.ld_g<PP>(g);
    if (sw) {
      float o[kK];
      seg_ldp(ox + PP * kPartBlk, cl, o); seg_stp(ox + PP * kPartBlk, cl, x + kK * PP);
#pragma unroll
      for (int j = 0; j < kK; ++j) x[kK * PP + j] = o[j];
      seg_ldp(om + PP * kPartBlk, cl, o); seg_stp(om + PP * kPartBlk, cl, m + kK * PP);
#pragma unroll
      for (int j = 0; j < kK; ++j) m[kK * PP + j] = o[j];
      seg_ldp(og + PP * kPartBlk, cl, o); seg_stp(og + PP * kPartBlk, cl, g);
#pragma unroll
      for (int j = 0; j < kK; ++j) g[j] = o[j];
    }
    if (doit) {
      seg_stp(bx + PP * kPartBlk, cl, x + kK * PP);
      seg_stp(bg + PP * kPartBlk, cl, g);
#pragma unroll
      for (int j = 0; j < kK; ++j) rho[kK * PP + j] = 0.f;
    }
    if (any_sw) cx.st_g<PP>(g);
  });
  if (any_sw) Ctx::wait_st();
}

// ---------------------------------------------------------------------------------------------
// tile128_nuts_kernel: lock-step.  L2 scratch: other end, trajectory / subtree candidates, rho, checkpoint slots.
enum { kVOx = 0, kVOm, kVOg, kVCx, kVCg, kVBx, kVBg, kVRho, kVCk };   // checkpoints: kVCk + slot (m), + depth + slot (rho)

__global__ void __launch_bounds__(kThreads, 1)
tile128_nuts_kernel(const ChainParams p, const DenseGaussianParams tp, float* __restrict__ scratch_all) {
  extern __shared__ __align__(128) unsigned char planes[];
  __shared__ Shared sh;
  Ctx cx;
  float* const dyn = reinterpret_cast<float*>(planes + 2 * kPlaneBytes);
  cx.init(&sh, planes, tp.P, tp.loc, tp.D, tp.scale);
  if (threadIdx.x >= kWorkers) {
    // third warpgroup: its first warp issues the contractions, the other three only exist so that the launch's
    // register allocation (384 x 168) covers what the workers take below
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if ((threadIdx.x >> 5) == kMmaWarp) cx.mma_loop();
    cx.finish();
    return;
  }
  // registers move inside the CTA's launch allocation: the third warpgroup hands its share back and the workers take
  // the 232 that keep x, m, rho of 52 dims in registers
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  Prof pf;
  pf.init();
  const int D = tp.D;
  const int cl = cx.cl;
  const int nvec = kVCk + 2 * p.max_depth;
  float* const scr_s = scratch_all + (size_t)blockIdx.x * nvec * kVS + (size_t)(kPT * cx.hf) * kPartBlk;
  auto sv = [&](int v) -> float* { return scr_s + (size_t)v * kVS; };
  LeafEnv e;
  e.lc = sh.loc + kDT * cx.hf;
  e.bx = sv(kVBx); e.bg = sv(kVBg); e.ck_m = sv(kVCk); e.ck_r = sv(kVCk + p.max_depth);
  e.ckl = dyn + (size_t)(kPT * cx.hf) * kPartBlk;
  e.hi_m = nullptr; e.hi_r = nullptr;
  e.flags = sh.flags;
  e.unrolled = p.unrolled; e.lognorm = tp.lognorm; e.max_energy_diff = p.max_energy_diff;
  e.Bg = (uint64_t)p.B_global; e.layout = p.layout;
  const int ntiles = (p.B + kM - 1) / kM;
  unsigned gt = 0;   // global leaf counter (rotates the "somebody continues" flags)
  for (int tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    const int c = tile_i * kM + cl;
    const bool live = c < p.B;
    const uint64_t cg = (uint64_t)p.chain_offset + (uint64_t)c;
    float x[kDT], m[kDT], rho[kDT];
    tile_load(p.x, c, D, cx.hf, live, x);
    {
      float g[kDT];
      tile_load(p.g, c, D, cx.hf, live, g);
      for_parts([&](auto pp) {
        constexpr int PP = decltype(pp)::value;
        cx.st_g<PP>(reinterpret_cast<const float(&)[kK]>(g[kK * PP]));
      });
      Ctx::wait_st();
    }
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
      // the tile's draws are spread over the whole CTA and handed over through shared memory (the previous-leaf
      // checkpoint buffer is dead between transitions)
      float* const mom = dyn;   // [kKP][kM]
      for (int w = threadIdx.x; w < kM * D; w += kWorkers) {
        const int li = w / D, d = w - li * D;
        const int cc = tile_i * kM + li;
        mom[d * kM + li] = cc < p.B ? nuts_momentum(p, sk, (uint64_t)p.chain_offset + (uint64_t)cc, d) : 0.f;
      }
      Ctx::wsync();
      float mm2[kPT];
      for_parts([&](auto pp) {
        constexpr int PP = decltype(pp)::value;
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const int d = kDT * cx.hf + kK * PP + j;
          const float mm = d < D ? mom[d * kM + cl] : 0.f;
          m[kK * PP + j] = mm;
          a = fmaf(mm, mm, a);
        }
        mm2[PP] = a;
        float g[kK];
        cx.ld_g<PP>(g);
        seg_stp(sv(kVOg) + PP * kPartBlk, cl, g);
        seg_stp(sv(kVCg) + PP * kPartBlk, cl, g);
      });
      seg_stv(sv(kVOx), cl, x); seg_stv(sv(kVOm), cl, m);
      seg_stv(sv(kVCx), cl, x);
      seg_stv(sv(kVRho), cl, m);
      float s1[1] = {sum4(mm2)};
      cx.reduce<1>(s1);
      const float H0 = lp - 0.5f * s1[0];
      float slp = lp, olp = lp, clp = lp, cen = H0, cw = 0.f;
      float esum = 0.f;
      int nleap = 0;
      bool cont = live, notdiv = true, accepted = false, s_is_right = true;
      int any_cont = Ctx::wsync_or(cont ? 1 : 0);
#pragma unroll 1
      for (int it = 0; it < p.max_depth && any_cont; ++it) {
        // per-depth randoms of this chain (nuts.py:551-558, :622-625)
        Key kd{hdr[6 * it], hdr[6 * it + 1]}, kac{hdr[6 * it + 2], hdr[6 * it + 3]};
        const bool dir = (bits_at(kd, cg, (uint64_t)p.B_global, p.layout) & 1u) != 0;
        const float lacc = log1pf(-uniform_from_bits(bits_at(kac, cg, (uint64_t)p.B_global, p.layout), 0.f, 1.f));
        const bool sw = dir != s_is_right;   // registers / D must hold the end that is extended
        // _build_sub_tree init (nuts.py:713-791)
        swap_and_seed(cx, true, sw, sv(kVOx), sv(kVOm), sv(kVOg), e.bx, e.bg, x, m, rho);
        if (sw) {
          const float a = slp; slp = olp; olp = a;
          s_is_right = dir;
        }
        const float eps = dir ? eps_abs : -eps_abs;
        LaneSub st;
        st.slp = slp; st.blp = slp; st.ben = slp; st.bw = -INFINITY; st.esum_sub = 0.f; st.n = 0;
        st.alive = cont; st.nd = notdiv;
        const int nsteps = 1 << it;
        const uint32_t* kud = ku + 2 * (nsteps - 1);
#pragma unroll 1
        for (int i = 0; i < nsteps; ++i, ++gt) {
          pf.mark(0);
          if (!nuts_leaf<true>(cx, e, i, gt, st.alive, 31, -1, 0, 0, 0, eps, H0, kud + 2 * i, cg, x, m, rho, st, pf)) break;
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
          float o[kDT];
          seg_ldv(e.bx, cl, o); seg_stv(sv(kVCx), cl, o);
          seg_ldv(e.bg, cl, o); seg_stv(sv(kVCg), cl, o);
          clp = st.blp; cen = st.ben;
        }
        float u0[kPT], u1[kPT];
        for_parts([&](auto pp) {
          constexpr int PP = decltype(pp)::value;
          float a0 = 0.f, a1 = 0.f;
          if (cont_f) {
            float rh[kK], om[kK];
            seg_ldp(sv(kVRho) + PP * kPartBlk, cl, rh);
            seg_ldp(sv(kVOm) + PP * kPartBlk, cl, om);
#pragma unroll
            for (int j = 0; j < kK; ++j) {
              const float rr = rh[j] + rho[kK * PP + j];
              rh[j] = rr;
              a0 = fmaf(rr, m[kK * PP + j], a0);
              a1 = fmaf(rr, om[j], a1);
            }
            seg_stp(sv(kVRho) + PP * kPartBlk, cl, rh);
          }
          u0[PP] = a0; u1[PP] = a1;
        });
        float s2[2] = {sum4(u0), sum4(u1)};
        cx.reduce<2>(s2);
        nleap += st.n;
        accepted = accepted || swap;
        notdiv = st.nd;
        cont = cont_f && (s2[0] >= 0.f) && (s2[1] >= 0.f);
        any_cont = Ctx::wsync_or(cont ? 1 : 0);       // nuts.py:404-407
      }
      // ---- results (nuts.py:424-445); the next state is the trajectory candidate
      seg_ldv(sv(kVCx), cl, x);
      float g[kDT];
      seg_ldv(sv(kVCg), cl, g);
      for_parts([&](auto pp) {
        constexpr int PP = decltype(pp)::value;
        cx.st_g<PP>(reinterpret_cast<const float(&)[kK]>(g[kK * PP]));
      });
      Ctx::wait_st();
      lp = clp;
      const int leap = nleap * p.unrolled;
      nleap_total += (unsigned long long)leap;
      const float lar = logf(esum / (float)nleap);
      if (live && cx.hf == 0 && p.lar_last) p.lar_last[c] = lar;
      if (r >= 0) {
        const Trace& tr = p.tr;
        if (tr.states) tile_store(tr.states, r, p.B, c, D, cx.hf, live, x);
        if (tr.grads) tile_store(tr.grads, r, p.B, c, D, cx.hf, live, g);
        if (live && cx.hf == 0) {
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
      if (t + 1 == p.t1) {
        tile_store(p.x, 0, p.B, c, D, cx.hf, live, x);
        tile_store(p.g, 0, p.B, c, D, cx.hf, live, g);
      }
    }
    if (live && cx.hf == 0) {
      p.lp[c] = lp;
      if (p.leapfrog_total) p.leapfrog_total[c] += nleap_total;
    }
  }
  cx.stop_mma();
  cx.finish();
}

// ---------------------------------------------------------------------------------------------
// tile128_nuts_async_kernel: every lane of the tile at its OWN position of its OWN tree (and its own transition) --
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
// idles until the end of its chunk (< 32 ticks).  After every transition a chain goes back to a ticket FIFO and the
// lane takes the chain at the head; every byte of a lane's in-flight state is private to the lane (registers, its
// TMEM lane, its columns of the CTA's scratch).
constexpr int kS0 = 5;                       // doublings of the START state; chunks are 2^kS0 ticks
constexpr int kChunkTicks = 1 << kS0;
enum { kLaneNone = 0, kLaneStart = 1, kLaneChunk = 2 };
// CTA scratch vectors: other end, trajectory candidate, rho, subtree candidate, 5 local checkpoint slots (m, rho),
// max_depth - 5 per-lane slots of chunk-first leaves (m, rho)
enum { kAOx = 0, kAOm, kAOg, kACx, kACg, kARho, kABx, kABg, kACkM, kACkR = kACkM + kS0, kAHiM = kACkR + kS0 };

// FIFO of chains waiting for their next transition (ring buffer of chain ids, -1 = empty slot): a lane hands its
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

__global__ void tile128_async_init_kernel(AsyncQueue aq, int B, int t0) {
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
tile128_nuts_async_kernel(const ChainParams p, const DenseGaussianParams tp, float* __restrict__ scratch_all,
                          const AsyncQueue aq) {
  extern __shared__ __align__(128) unsigned char planes[];
  __shared__ Shared sh;
  __shared__ int new_chain[kM];
  __shared__ int hi_max, n_start;
  __shared__ int st_lane[kM], st_c[kM], st_t[kM];
  Ctx cx;
  float* const dyn = reinterpret_cast<float*>(planes + 2 * kPlaneBytes);
  cx.init(&sh, planes, tp.P, tp.loc, tp.D, tp.scale);
  if (threadIdx.x >= kWorkers) {
    // third warpgroup: its first warp issues the contractions, the other three only exist so that the launch's
    // register allocation (384 x 168) covers what the workers take below
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if ((threadIdx.x >> 5) == kMmaWarp) cx.mma_loop();
    cx.finish();
    return;
  }
  // registers move inside the CTA's launch allocation: the third warpgroup hands its share back and the workers take
  // the 232 that keep x, m, rho of 52 dims in registers
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  Prof pf;
  pf.init();
  const int D = tp.D;
  const int cl = cx.cl;
  const int nhi = p.max_depth > kS0 ? p.max_depth - kS0 : 0;
  const int nvec = kAHiM + 2 * nhi;
  float* const scr_s = scratch_all + (size_t)blockIdx.x * nvec * kVS + (size_t)(kPT * cx.hf) * kPartBlk;
  auto sv = [&](int v) -> float* { return scr_s + (size_t)v * kVS; };
  LeafEnv e;
  e.lc = sh.loc + kDT * cx.hf;
  e.bx = sv(kABx); e.bg = sv(kABg); e.ck_m = sv(kACkM); e.ck_r = sv(kACkR);
  e.ckl = dyn + (size_t)(kPT * cx.hf) * kPartBlk;
  e.hi_m = sv(kAHiM); e.hi_r = sv(kAHiM + nhi);
  e.flags = nullptr;
  e.unrolled = p.unrolled; e.lognorm = tp.lognorm; e.max_energy_diff = p.max_energy_diff;
  const uint64_t Bg = (uint64_t)p.B_global;
  e.Bg = Bg; e.layout = p.layout;

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
  float x[kDT], m[kDT], rho[kDT];
#pragma unroll
  for (int j = 0; j < kDT; ++j) { x[j] = 0.f; m[j] = 0.f; rho[j] = 0.f; }
  const unsigned long long qcap = (unsigned long long)aq.cap;
  const long long total_pushes = (long long)p.B * (long long)(p.t1 - p.t0);
  long long ticket = -1;   // (half 0 of a lane) my position in the FIFO while the lane waits for a chain

  // per-transition key schedule of this lane
  auto keys = [&](const uint32_t*& sk, const uint32_t*& hdr, const uint32_t*& ku) {
    sk = p.sched + (size_t)(t - p.t_sched0) * p.sched_stride;
    hdr = sk + 2 * p.n_parts;
    ku = hdr + 6 * p.max_depth;
  };
  auto direction = [&]() -> bool {   // nuts.py:551-558
    const uint32_t *sk, *hdr, *ku;
    keys(sk, hdr, ku);
    Key kd{hdr[6 * it], hdr[6 * it + 1]};
    return (bits_at(kd, (uint64_t)p.chain_offset + (uint64_t)c, Bg, p.layout) & 1u) != 0;
  };
  auto seed_subtree = [&](bool dir) {   // _build_sub_tree init (nuts.py:713-791), scalars
    st.blp = st.slp; st.ben = st.slp; st.bw = -INFINITY; st.esum_sub = 0.f; st.n = 0; st.nd = notdiv; st.alive = true;
    ihi = 0;
    eps = dir ? eps_abs : -eps_abs;
  };

  // ---- start transition t of chain c from (x, g, lp): _start_trajectory_batched (nuts.py:512-539).  Tile-uniform
  // (one pair exchange); `go` selects the lanes that start.  The gradient does not go to D here: tick 0 of the chunk
  // (an idle tick for START lanes) runs a contraction over it; the lanes fetch it from the other-end copy after
  // that tick.
  auto start_transitions = [&](bool go) {
    // momentum ~ N(0, I).  After the first chunk only a few lanes start at a time: their draws are spread over the
    // whole CTA and handed over through shared memory (the previous-leaf checkpoint buffer is dead between chunks)
    if (threadIdx.x == 0) n_start = 0;
    Ctx::wsync();
    if (go && cx.hf == 0) {
      const int idx = atomicAdd(&n_start, 1);
      st_lane[idx] = cl; st_c[idx] = c; st_t[idx] = t;
    }
    Ctx::wsync();
    float* const mom = dyn;   // [kKP][kM]
    for (int w = threadIdx.x; w < n_start * D; w += kWorkers) {
      const int li = w / D, d = w - li * D;
      const uint32_t* skl = p.sched + (size_t)(st_t[li] - p.t_sched0) * p.sched_stride;
      mom[d * kM + st_lane[li]] = nuts_momentum(p, skl, (uint64_t)p.chain_offset + (uint64_t)st_c[li], d);
    }
    Ctx::wsync();
    pf.mark(12);
    float mm2[kPT] = {0.f, 0.f, 0.f, 0.f};
    if (go) {
      for_parts([&](auto pp) {
        constexpr int PP = decltype(pp)::value;
        float a = 0.f;
        float g[kK];
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const int d = kDT * cx.hf + kK * PP + j;
          const float mm = d < D ? mom[d * kM + cl] : 0.f;
          m[kK * PP + j] = mm;
          a = fmaf(mm, mm, a);
          g[j] = d < D ? __ldcg(p.g + (size_t)c * D + d) : 0.f;
          rho[kK * PP + j] = 0.f;
        }
        mm2[PP] = a;
        seg_stp(sv(kAOg) + PP * kPartBlk, cl, g);
        seg_stp(sv(kACg) + PP * kPartBlk, cl, g);
        seg_stp(e.bg + PP * kPartBlk, cl, g);
      });
      seg_stv(sv(kAOx), cl, x); seg_stv(sv(kAOm), cl, m);
      seg_stv(sv(kACx), cl, x);
      seg_stv(sv(kARho), cl, m);
      seg_stv(e.bx, cl, x);
    }
    float s1[1] = {sum4(mm2)};
    cx.reduce<1>(s1);
    if (go) {
      H0 = lp - 0.5f * s1[0];
      st.slp = lp; olp = lp; clp = lp; cen = H0; cw = 0.f; esum = 0.f;
      nleap = 0;
      cont = true; notdiv = true; accepted = false;
      type = kLaneStart; fin = false; it = 0;
      // both ends are the initial point: extending "the other end" needs no exchange
      const bool dir = direction();
      s_is_right = dir;
      seed_subtree(dir);
    }
  };

  // ---- end the current doubling of the selected lanes (_loop_tree_doubling tail, nuts.py:597-711) and begin the
  // next one, or mark the transition finished.  Tile-uniform (one pair exchange).
  auto doubling_boundary = [&](bool endd, bool chunk_end) {
    float u0[kPT] = {0.f, 0.f, 0.f, 0.f}, u1[kPT] = {0.f, 0.f, 0.f, 0.f};
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
        float o[kDT];
        seg_ldv(e.bx, cl, o); seg_stv(sv(kACx), cl, o);
        seg_ldv(e.bg, cl, o); seg_stv(sv(kACg), cl, o);
        clp = st.blp; cen = st.ben;
      }
      if (st.alive) {
        for_parts([&](auto pp) {
          constexpr int PP = decltype(pp)::value;
          float a0 = 0.f, a1 = 0.f;
          float rh[kK], om[kK];
          seg_ldp(sv(kARho) + PP * kPartBlk, cl, rh);
          seg_ldp(sv(kAOm) + PP * kPartBlk, cl, om);
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            const float rr = rh[j] + rho[kK * PP + j];
            rh[j] = rr;
            a0 = fmaf(rr, m[kK * PP + j], a0);
            a1 = fmaf(rr, om[j], a1);
          }
          seg_stp(sv(kARho) + PP * kPartBlk, cl, rh);
          u0[PP] = a0; u1[PP] = a1;
        });
      }
    }
    float s2[2] = {sum4(u0), sum4(u1)};
    cx.reduce<2>(s2);
    bool next = false, dir = false, sw = false;
    if (endd) {
      nleap += st.n;
      accepted = accepted || swap;
      notdiv = st.nd;
      cont = st.alive && (s2[0] >= 0.f) && (s2[1] >= 0.f);
      if (cont && it + 1 < p.max_depth) {
        it += 1;
        if (chunk_end && type == kLaneStart) type = kLaneChunk;
        next = true;
        dir = direction();
        sw = dir != s_is_right;
      } else {
        fin = true;
        st.alive = false;
      }
    }
    swap_and_seed(cx, next, sw, sv(kAOx), sv(kAOm), sv(kAOg), e.bx, e.bg, x, m, rho);
    if (next) {
      if (sw) {
        const float a = st.slp; st.slp = olp; olp = a;
        s_is_right = dir;
      }
      seed_subtree(dir);
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
      if (want && cx.hf == 0) {
        if (ticket < 0) ticket = (long long)atomicAdd(aq.ctl + kQHead, 1ull);
        int nc = -1;
        if (ticket < total_pushes) {
          volatile int* slot = aq.q + (size_t)((unsigned long long)ticket % qcap);
          nc = *slot;
          if (nc >= 0) { *slot = -1; ticket = -1; }
        }
        new_chain[cl] = nc;
      }
      Ctx::wsync();   // (the chain's state is read with ld.cg from L2, where its last owner's fenced stores are)
      bool go = false;
      if (want) {
        c = new_chain[cl];
        if (c >= 0) {
#pragma unroll
          for (int j = 0; j < kDT; ++j) {
            const bool in = kDT * cx.hf + j < D;
            x[j] = in ? __ldcg(p.x + (size_t)c * D + kDT * cx.hf + j) : 0.f;
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
    const int has_start = Ctx::wsync_or(type == kLaneStart ? 1 : 0);
    const int any_work = Ctx::wsync_or(type != kLaneNone ? 1 : 0);
    if (!any_work) {
      // lanes that still hold a servable ticket wait for their chain; otherwise every lane of the tile has retired
      const int pending = Ctx::wsync_or(cx.hf == 0 && ticket >= 0 && ticket < total_pushes ? 1 : 0);
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
    if (cx.hf == 0 && t_hi > 0) atomicMax(&hi_max, t_hi);
    Ctx::wsync();
    const int hi_checks = hi_max;

#pragma unroll 1
    for (int i = 0; i < kChunkTicks; ++i) {
      pf.mark(0);
      const bool act = st.alive && !fin && (type == kLaneChunk || (type == kLaneStart && i >= 1));
      const bool has_key = type == kLaneChunk || (type == kLaneStart && i >= 1);
      // the largest subtree (2^jmax leaves) that can close inside this lane's doubling at this tick
      const int jmax = type == kLaneChunk ? kS0 : 31 - __clz(i | 1);
      nuts_leaf<false>(cx, e, i, 0u, act, jmax, hi_slot_w, ihi, t_hi, i == kChunkTicks - 1 ? hi_checks : 0, eps, H0,
                       has_key ? kb + 2 * i : nullptr, cg, x, m, rho, st, pf);
      if (i == 0 && has_start) {
        // START lanes: the idle tick's contraction ran over D; fetch the gradient of the initial point
        const bool rs = type == kLaneStart && !fin;
        if (__any_sync(kFull, rs)) {
          for_parts([&](auto pp) {
            constexpr int PP = decltype(pp)::value;
            float g[kK];
            cx.ld_g<PP>(g);
            if (rs) seg_ldp(sv(kAOg) + PP * kPartBlk, cl, g);
            cx.st_g<PP>(g);
          });
          Ctx::wait_st();
        }
      }
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
        float g[kDT];
        seg_ldv(sv(kACx), cl, x);
        seg_ldv(sv(kACg), cl, g);
        lp = clp;
        const int leap = nleap * p.unrolled;
        const float lar = logf(esum / (float)nleap);
        if (cx.hf == 0 && p.lar_last) p.lar_last[c] = lar;
        const int r = nuts_result_index(p, t);
        if (r >= 0) {
          const Trace& tr = p.tr;
          if (tr.states) tile_store(tr.states, r, p.B, c, D, cx.hf, true, x);
          if (tr.grads) tile_store(tr.grads, r, p.B, c, D, cx.hf, true, g);
          if (cx.hf == 0) {
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
        tile_store(p.x, 0, p.B, c, D, cx.hf, true, x);
        tile_store(p.g, 0, p.B, c, D, cx.hf, true, g);
        if (cx.hf == 0) {
          p.lp[c] = lp;
          aq.t_next[c] = t;
          if (p.leapfrog_total) atomicAdd(p.leapfrog_total + c, (unsigned long long)leap);
        }
        type = kLaneNone; fin = false; st.alive = false;
      }
      __threadfence();
      Ctx::wsync();   // both halves of every finished chain have written its state
      if (done && cx.hf == 0 && t < p.t1) {
        const unsigned long long sl = atomicAdd(aq.ctl + kQTail, 1ull);
        *(volatile int*)(aq.q + (size_t)(sl % qcap)) = c;
      }
    }
    pf.mark(9);
  }
  cx.stop_mma();
  cx.finish();
}

#ifdef PB2_TILE_PROF
static void dump_tile_prof(pb2_ctx* ctx) {
  unsigned long long h[2][16];
  cudaStreamSynchronize(ctx->stream);
  cudaMemcpyFromSymbol(h, g_tile_prof, sizeof(h));
  static const char* nm[13] = {"head", "kick+stage", "contract", "post", "exchange", "extra checks", "scalars+take",
                               "START boundaries", "chunk-end boundary", "finish", "start(rest)", "acquire", "momentum"};
  for (int w = 0; w < 2; ++w) {
    fprintf(stderr, "[tile128prof t%d] leaves %llu:", w ? kWorkers - 1 : 0, h[w][15]);
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

}  // namespace t128

int launch_tile128_nuts(pb2_ctx* ctx, const pb2_target* tgt, ChainParams& p, bool lockstep) {
  using namespace t128;
  DenseGaussianParams tp{tgt->d_a, tgt->d_b, tgt->scalar, tgt->dim};
  tp.scale = p.scale;
  // P hi/lo planes + the previous leaf's checkpoint (momentum, rho)
  const size_t smem = 2 * (size_t)kPlaneBytes + 2 * kVS * sizeof(float);
  const int ntiles = (p.B + kM - 1) / kM;
  if (!lockstep) {
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
    tile128_async_init_kernel<<<(aq.cap + 255) / 256, 256, 0, ctx->stream>>>(aq, p.B, p.t0);
    if (int rc = check_cuda(ctx, cudaFuncSetAttribute(tile128_nuts_async_kernel,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(tile128_nuts_async)"))
      return rc;
    tile128_nuts_async_kernel<<<agrid, kThreads, smem, ctx->stream>>>(p, tp, ctx->d_ckpt, aq);
    ctx->launches += 2;
    dump_tile_prof(ctx);
    return check_cuda(ctx, cudaGetLastError(), "tile128_nuts_async_kernel");
  }
  const int grid = std::min(ntiles, ctx->num_sms);
  const size_t need = (size_t)(kVCk + 2 * p.max_depth) * kVS * sizeof(float) * grid;
  if (int rc = ensure_scratch(ctx, need, "cudaMalloc(tile scratch)")) return rc;
  if (int rc = check_cuda(ctx, cudaFuncSetAttribute(tile128_nuts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)smem), "cudaFuncSetAttribute(tile128_nuts)"))
    return rc;
  tile128_nuts_kernel<<<grid, kThreads, smem, ctx->stream>>>(p, tp, ctx->d_ckpt);
  ctx->launches += 1;
  dump_tile_prof(ctx);
  return check_cuda(ctx, cudaGetLastError(), "tile128_nuts_kernel");
}

}  // namespace pb2
