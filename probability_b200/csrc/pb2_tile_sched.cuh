// tile_nuts_sched_kernel: the tcgen05 tile NUTS kernel with chains re-grouped every 32 leaves.
//
// The lock-step tile kernel (tile_nuts_kernel) makes all 128 chains of a tile wait for the tile's
// deepest tree (measured utilisation 0.27 on the 100-d ill-conditioned Gaussian: mean 276 of max
// 1023 leapfrogs).  Here the unit of work is a TASK of at most 32 leaves, and there are only two kinds:
//   class 0 (START): start a transition (momentum, H0, ...) and run doublings 0 .. 4 (31 leaves)
//   class 1 (CHUNK): 32 consecutive leaves [32 q, 32 q + 32) of a doubling it >= 5 -- every chain of the
//                    tile at its OWN doubling `it` and its OWN chunk q.  That works in lock-step because the
//                    leaf index within the chunk (the low 5 bits of the leaf number) decides everything that
//                    must be uniform: checkpoint write vs U-turn check, the popcount slot, how many of the
//                    2-, 4-, .., 32-leaf subtrees close.  Only the last leaf's checks of the 64-, 128-, ..
//                    leaf subtrees differ per chain; they read per-chain checkpoints from the chain's record.
// Persistent CTAs pop up to 128 chain ids from the class's queue, run the task exactly like
// tile_nuts_kernel, and push every chain to the queue of its next task (or finish its transition,
// emit the traced results and push it to class 0).  A chain that U-turns inside a subtree leaves at
// the end of its chunk, so a dead chain rides along for < 32 leaves, and with two classes there is
// (almost) always a full tile to run.
// A chain's results do not depend on which tile / CTA ran its tasks: every random number is a function
// of (step seed, global chain index) and all per-chain reductions have a fixed order -- the kernel is
// bit-identical to tile_nuts_kernel (tests/test_gpu_parity.py).
// Chain-major records hold what survives a task boundary: both trajectory ends, the trajectory
// candidate, rho, the subtree candidate, rho_subtree and the checkpoints of chunk-first leaves.
#pragma once
#include "pb2_tile_nuts.cuh"

namespace pb2 {
using namespace tile;

enum { kRSx = 0, kRSm, kRSg, kROx, kROm, kROg, kRCx, kRCg, kRRho, kRBx, kRBg, kRRs, kRHi };   // + 2 * (max_depth - 5)
enum {
  kSLp = 0, kSH0, kSSlp, kSOlp, kSClp, kSCen, kSCw, kSEsum, kSNleap, kSFlags, kST, kSIt, kSIhi,
  kSBlp, kSBen, kSBw, kSEsub, kSN, kRecScal = 32
};
constexpr int kS0 = 5;            // doublings merged into the START task; chunks are 2^kS0 leaves
constexpr int kChunkLeaves = 1 << kS0;
enum { kQHead0 = 0, kQTail0, kQHead1, kQTail1, kQDone, kQRunning, kQWords = 8 };

struct SchedParams {
  float* rec_vec;    // [B][nrv][kRecStride]
  float* rec_scal;   // [B][kRecScal]
  int* queue;        // [2][B] ring buffers of chain ids (-1 = empty slot)
  unsigned long long* qctl;   // [kQWords] heads / tails / finished chains / running chains
  int nrv;           // vectors per record: kRHi + 2 (max_depth - 5)
  int patience;      // polls without a full tile before a partial tile is accepted
  unsigned long long* stats;   // optional [32]: tasks, claimed chains, ticks, idle polls, per-class tasks @8+, cycles @26+
};

__global__ void tile_sched_init_kernel(SchedParams sp, int B, int t0) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) {
    for (int k = 0; k < kQWords; ++k) sp.qctl[k] = 0ull;
    sp.qctl[kQTail0] = (unsigned long long)B;
  }
  if (c >= B) return;
  sp.queue[c] = c;
  sp.queue[B + c] = -1;
  sp.rec_scal[(size_t)c * kRecScal + kST] = __int_as_float(t0);
}

__global__ void __launch_bounds__(kThreads, 1)
tile_nuts_sched_kernel(const ChainParams p, const DenseGaussianParams tp, const SchedParams sp,
                       float* __restrict__ scratch_all) {
  extern __shared__ __align__(128) unsigned char planes[];
  __shared__ Shared sh;
  __shared__ float lu[4][kM];
  __shared__ int ids[kM];
  __shared__ int n_claimed, sel_cls, hi_max;
  __shared__ unsigned long long q_head;
  Ctx cx;
  cx.init(&sh, planes, tp.P, tp.loc, tp.D);
  Prof pf;
  pf.init();
  const int D = tp.D;
  const int tid = threadIdx.x;
  const int cl = cx.cl;
  // CTA scratch: subtree candidate (x, g) + 5 local checkpoint slots; per-thread segments (pb2_tile.cuh seg_*)
  enum { kTBx = 0, kTBg = 1, kTCk = 2, kTVecs = kTCk + 2 * kS0 };
  float* const scr_s = scratch_all + (size_t)blockIdx.x * kTVecs * kVS + (size_t)(kK * cx.slice) * kM;
  SubtreeArgs sa;
  sa.unrolled = p.unrolled; sa.layout = p.layout; sa.b_global = (uint64_t)p.B_global;
  sa.lognorm = tp.lognorm; sa.max_energy_diff = p.max_energy_diff;
  sa.lc = sh.loc + kK * cx.slice;
  sa.bx = scr_s + (size_t)kTBx * kVS; sa.bg = scr_s + (size_t)kTBg * kVS;
  sa.ck_m = scr_s + (size_t)kTCk * kVS; sa.ck_r = scr_s + (size_t)(kTCk + kS0) * kVS;
  sa.ckl = reinterpret_cast<float*>(planes + 2 * kPlaneBytes) + (size_t)(kK * cx.slice) * kM;
  sa.hck = nullptr; sa.ihi = 0; sa.hi_slot_w = -1; sa.hi_checks = 0;
  unsigned gt = 0;
  const unsigned long long Bq = (unsigned long long)p.B;
  long long tmark = clock64();
  auto lap = [&](int k) {   // thread 0 of every CTA: cycles per scheduler phase (only with PB2_SCHED_STATS)
    if (sp.stats && tid == 0) {
      const long long now = clock64();
      atomicAdd(sp.stats + k, (unsigned long long)(now - tmark));
      tmark = now;
    }
  };

  while (true) {
    // ------------------------------------------------------------ (1) pop up to 128 chains of one class
    __syncthreads();   // everybody is done with the shared scheduling words of the previous task
    if (tid == 0) {
      volatile unsigned long long* q = sp.qctl;
      int cls = -1, n = 0, patience = 0;
      unsigned long long hsel = 0;
      while (true) {
        if (q[kQDone] >= Bq) { cls = -2; break; }
        const unsigned long long h0 = q[kQHead0], h1 = q[kQHead1];
        const long long a0 = (long long)(q[kQTail0] - h0), a1 = (long long)(q[kQTail1] - h1);
        // a full tile first (of the class with more waiting chains); a partial one when producers stay away
        // for a while, or immediately when nothing is running anywhere (nobody will push more)
        const int pick = a1 >= a0 ? 1 : 0;
        const long long av = pick ? a1 : a0;
        if (av >= kM || (av > 0 && (patience >= sp.patience || q[kQRunning] == 0ull))) {
          n = (int)(av < kM ? av : kM);
          const unsigned long long h = pick ? h1 : h0;
          if (atomicCAS(sp.qctl + (pick ? kQHead1 : kQHead0), h, h + (unsigned long long)n) == h) {
            atomicAdd(sp.qctl + kQRunning, (unsigned long long)n);
            cls = pick; hsel = h;
            break;
          }
          continue;
        }
        if (sp.stats) atomicAdd(sp.stats + 3, 1ull);
        patience++;
        __nanosleep(2000);
      }
      sel_cls = cls; n_claimed = n; q_head = hsel; hi_max = 0;
    }
    __syncthreads();
    const int cls = sel_cls;
    if (cls == -2) break;
    const int ntask = n_claimed;
    if (tid < ntask) {
      volatile int* slot = sp.queue + (size_t)cls * p.B + (size_t)((q_head + (unsigned long long)tid) % Bq);
      int id;
      while ((id = *slot) < 0) {}   // the pusher reserved the slot before it wrote the id
      *slot = -1;
      ids[tid] = id;
    }
    __threadfence();
    __syncthreads();
    lap(26);
    if (sp.stats && tid == 0) {
      atomicAdd(sp.stats + 0, 1ull);
      atomicAdd(sp.stats + 1, (unsigned long long)ntask);
      atomicAdd(sp.stats + 8 + cls, 1ull);
      atomicAdd(sp.stats + 20 + cls, (unsigned long long)ntask);
    }
    const unsigned gt_begin = gt;
    // ------------------------------------------------------------ (2) load the chains' state, run the task
    const bool live = cl < ntask;
    const int c = live ? ids[cl] : 0;
    const uint64_t cg = (uint64_t)p.chain_offset + (uint64_t)c;
    sa.cg = cg;
    float* const rec = sp.rec_vec + (size_t)c * sp.nrv * kRecStride + kRecSlice * cx.slice;
    auto rv = [&](int v) -> float* { return rec + (size_t)v * kRecStride; };
    float* const rs = sp.rec_scal + (size_t)c * kRecScal;
    const int t = live ? __float_as_int(__ldcg(&rs[kST])) : p.t0;
    const float eps_abs = p.step_kind == 0 ? p.step[0] : (live ? p.step[c] : 0.f);
    const uint32_t* sk = p.sched + (size_t)(t - p.t_sched0) * p.sched_stride;
    const uint32_t* hdr = sk + 2 * p.n_parts;
    const uint32_t* ku = hdr + 6 * p.max_depth;
    float x[kK], m[kK], rho[kK];
    float lp, H0, slp, olp, clp, cen, cw, esum;
    int nleap;
    bool cont, notdiv, accepted, s_is_right;
    bool to_chunk = false;     // the chain's next task is a CHUNK (its state is saved below)
    int it_next = 0, ihi_next = 0;
    SubtreeState st;
    st.blp = st.ben = st.bw = st.esum_sub = st.slp = 0.f; st.n = 0; st.c_prev = false; st.nd = true; st.took = false;

    if (cls == 0) {
      // ---- START: _start_trajectory_batched (nuts.py:512-539), then doublings 0 .. kS0 - 1
      {
        float g[kK];
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const bool in = live && (kK * cx.slice + j < D);
          x[j] = in ? __ldcg(p.x + (size_t)c * D + kK * cx.slice + j) : 0.f;
          g[j] = in ? __ldcg(p.g + (size_t)c * D + kK * cx.slice + j) : 0.f;
        }
        cx.store_d(g);
        lp = live ? __ldcg(p.lp + c) : 0.f;
        float s1[1] = {0.f};
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const int d = kK * cx.slice + j;
          const float mm = (live && d < D) ? tile_momentum(p, sk, cg, d) : 0.f;
          m[j] = mm;
          s1[0] = fmaf(mm, mm, s1[0]);
        }
        if (live) {
          rec_st26(rv(kROx), x); rec_st26(rv(kROm), m); rec_st26(rv(kROg), g);
          rec_st26(rv(kRCx), x); rec_st26(rv(kRCg), g);
          rec_st26(rv(kRRho), m);
        }
        cx.reduce<1>(s1);
        H0 = lp - 0.5f * s1[0];
      }
      sa.H0 = H0;
      slp = lp; olp = lp; clp = lp; cen = H0; cw = 0.f; esum = 0.f;
      nleap = 0;
      cont = live; notdiv = true; accepted = false; s_is_right = true;
      int any_cont = __syncthreads_or(cont ? 1 : 0);
      lap(28);
      const int it_end = min(kS0, p.max_depth);
#pragma unroll 1
      for (int it = 0; it < it_end && any_cont; ++it) {
        Key kd{hdr[6 * it], hdr[6 * it + 1]}, kac{hdr[6 * it + 2], hdr[6 * it + 3]};
        const bool dir = (bits_at(kd, cg, (uint64_t)p.B_global, p.layout) & 1u) != 0;
        const float lacc = log1pf(-uniform_from_bits(bits_at(kac, cg, (uint64_t)p.B_global, p.layout), 0.f, 1.f));
        {
          const bool sw = live && dir != s_is_right;
          float g[kK];
          cx.load_d(g);
          if (sw) {
            float o[kK];
            rec_ld26(rv(kROx), o); rec_st26(rv(kROx), x);
#pragma unroll
            for (int j = 0; j < kK; ++j) x[j] = o[j];
            rec_ld26(rv(kROm), o); rec_st26(rv(kROm), m);
#pragma unroll
            for (int j = 0; j < kK; ++j) m[j] = o[j];
            rec_ld26(rv(kROg), o); rec_st26(rv(kROg), g);
#pragma unroll
            for (int j = 0; j < kK; ++j) g[j] = o[j];
            const float a = slp; slp = olp; olp = a;
            s_is_right = dir;
          }
          if (__any_sync(0xffffffffu, sw)) cx.store_d(g);
          seg_st26(sa.bx, cl, x);
          seg_st26(sa.bg, cl, g);
        }
        sa.eps = dir ? eps_abs : -eps_abs;
        sa.nsteps = 1 << it;
        sa.kud = ku + 2 * (sa.nsteps - 1);
        st.slp = slp; st.c_prev = cont; st.nd = notdiv;
        nuts_subtree<false>(cx, sh, lu, gt, sa, x, m, rho, st, pf);
        slp = st.slp;
        const bool cont_f = st.c_prev;
        esum = st.esum_sub + esum;
        const float tw = cont_f ? st.bw : -INFINITY;
        const float wsum = log_add_exp(tw, cw);
        float thr = tw - cw;
        thr = isnan(thr) ? 0.f : thr;
        const bool swap = (lacc <= thr) && cont_f;
        cw = wsum;
        if (swap && live) {
          float o[kK];
          seg_ld26(sa.bx, cl, o); rec_st26(rv(kRCx), o);
          seg_ld26(sa.bg, cl, o); rec_st26(rv(kRCg), o);
          clp = st.blp; cen = st.ben;
        }
        float s2[2] = {0.f, 0.f};
        if (live) {
          float rh[kK], om[kK];
          rec_ld26(rv(kRRho), rh);
          rec_ld26(rv(kROm), om);
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            const float rr = rh[j] + rho[j];
            rh[j] = rr;
            s2[0] = fmaf(rr, m[j], s2[0]);
            s2[1] = fmaf(rr, om[j], s2[1]);
          }
          rec_st26(rv(kRRho), rh);
        }
        cx.reduce<2>(s2);
        nleap += st.n;
        accepted = accepted || swap;
        notdiv = st.nd;
        cont = cont_f && (s2[0] >= 0.f) && (s2[1] >= 0.f);
        any_cont = __syncthreads_or(cont ? 1 : 0);
      }
      to_chunk = cont && it_end < p.max_depth;
      it_next = it_end; ihi_next = 0;
    } else {
      // ---- CHUNK: leaves [32 ihi, 32 ihi + 32) of this chain's doubling `it`
      int it = kS0, ihi = 0;
      bool bflag = false;        // the subtree candidate in the CTA scratch is this chain's current one
      {
        float g[kK];
        if (live) {
          rec_ld26(rv(kRSx), x); rec_ld26(rv(kRSm), m); rec_ld26(rv(kRSg), g);
          lp = __ldcg(&rs[kSLp]); H0 = __ldcg(&rs[kSH0]); slp = __ldcg(&rs[kSSlp]); olp = __ldcg(&rs[kSOlp]);
          clp = __ldcg(&rs[kSClp]); cen = __ldcg(&rs[kSCen]); cw = __ldcg(&rs[kSCw]); esum = __ldcg(&rs[kSEsum]);
          nleap = __float_as_int(__ldcg(&rs[kSNleap]));
          const int fl = __float_as_int(__ldcg(&rs[kSFlags]));
          notdiv = (fl & 2) != 0; accepted = (fl & 4) != 0; s_is_right = (fl & 8) != 0; st.nd = (fl & 16) != 0;
          it = __float_as_int(__ldcg(&rs[kSIt])); ihi = __float_as_int(__ldcg(&rs[kSIhi]));
        } else {
#pragma unroll
          for (int j = 0; j < kK; ++j) { x[j] = 0.f; m[j] = 0.f; g[j] = 0.f; }
          lp = H0 = slp = olp = clp = cen = cw = esum = 0.f;
          nleap = 0; notdiv = true; accepted = false; s_is_right = true;
        }
        cont = live;
        if (live && ihi == 0) {
          // a new doubling: direction (nuts.py:551-558), registers / D must hold the end that is extended,
          // _build_sub_tree init (nuts.py:713-791)
          Key kd{hdr[6 * it], hdr[6 * it + 1]};
          const bool dir = (bits_at(kd, cg, (uint64_t)p.B_global, p.layout) & 1u) != 0;
          if (dir != s_is_right) {
            float o[kK];
            rec_ld26(rv(kROx), o); rec_st26(rv(kROx), x);
#pragma unroll
            for (int j = 0; j < kK; ++j) x[j] = o[j];
            rec_ld26(rv(kROm), o); rec_st26(rv(kROm), m);
#pragma unroll
            for (int j = 0; j < kK; ++j) m[j] = o[j];
            rec_ld26(rv(kROg), o); rec_st26(rv(kROg), g);
#pragma unroll
            for (int j = 0; j < kK; ++j) g[j] = o[j];
            const float a = slp; slp = olp; olp = a;
            s_is_right = dir;
          }
          seg_st26(sa.bx, cl, x);
          seg_st26(sa.bg, cl, g);
          bflag = true;
#pragma unroll
          for (int j = 0; j < kK; ++j) rho[j] = 0.f;
          st.blp = slp; st.ben = slp; st.bw = -INFINITY; st.esum_sub = 0.f; st.n = 0; st.nd = notdiv;
        } else if (live) {
          rec_ld26(rv(kRRs), rho);
          st.blp = __ldcg(&rs[kSBlp]); st.ben = __ldcg(&rs[kSBen]); st.bw = __ldcg(&rs[kSBw]);
          st.esum_sub = __ldcg(&rs[kSEsub]); st.n = __float_as_int(__ldcg(&rs[kSN]));
        } else {
#pragma unroll
          for (int j = 0; j < kK; ++j) rho[j] = 0.f;
        }
        cx.store_d(g);
      }
      sa.H0 = H0;
      sa.eps = s_is_right ? eps_abs : -eps_abs;
      sa.nsteps = kChunkLeaves;
      sa.kud = ku + 2 * ((1 << it) - 1) + 2 * (ihi * kChunkLeaves);
      sa.hck = rv(kRHi);
      sa.ihi = ihi;
      const int nchunks = 1 << (it - kS0);
      sa.hi_slot_w = (live && (ihi & 1) == 0 && ihi + 1 < nchunks) ? __popc(ihi) : -1;
      if (live && cx.slice == 0) atomicMax(&hi_max, __ffs(~ihi) - 1);
      st.slp = slp; st.c_prev = cont;
      __syncthreads();
      sa.hi_checks = hi_max;
      lap(28);
      nuts_subtree<true>(cx, sh, lu, gt, sa, x, m, rho, st, pf);
      slp = st.slp;
      const bool cont_f = st.c_prev;                                   // no U-turn / divergence inside the subtree so far
      const bool end_doubling = live && (!cont_f || ihi + 1 == nchunks);
      const bool b_in_scratch = bflag || st.took;
      float s2[2] = {0.f, 0.f};
      bool swap = false;
      if (end_doubling) {
        // _loop_tree_doubling tail (nuts.py:597-711)
        Key kac{hdr[6 * it + 2], hdr[6 * it + 3]};
        const float lacc = log1pf(-uniform_from_bits(bits_at(kac, cg, (uint64_t)p.B_global, p.layout), 0.f, 1.f));
        esum = st.esum_sub + esum;
        const float tw = cont_f ? st.bw : -INFINITY;
        const float wsum = log_add_exp(tw, cw);
        float thr = tw - cw;
        thr = isnan(thr) ? 0.f : thr;
        swap = (lacc <= thr) && cont_f;
        cw = wsum;
        if (swap) {
          float o[kK];
          if (b_in_scratch) {
            seg_ld26(sa.bx, cl, o); rec_st26(rv(kRCx), o);
            seg_ld26(sa.bg, cl, o); rec_st26(rv(kRCg), o);
          } else {
            rec_ld26(rv(kRBx), o); rec_st26(rv(kRCx), o);
            rec_ld26(rv(kRBg), o); rec_st26(rv(kRCg), o);
          }
          clp = st.blp; cen = st.ben;
        }
        if (cont_f) {
          float rh[kK], om[kK];
          rec_ld26(rv(kRRho), rh);
          rec_ld26(rv(kROm), om);
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            const float rr = rh[j] + rho[j];
            rh[j] = rr;
            s2[0] = fmaf(rr, m[j], s2[0]);
            s2[1] = fmaf(rr, om[j], s2[1]);
          }
          rec_st26(rv(kRRho), rh);
        }
      } else if (live && b_in_scratch) {
        float o[kK];
        seg_ld26(sa.bx, cl, o); rec_st26(rv(kRBx), o);
        seg_ld26(sa.bg, cl, o); rec_st26(rv(kRBg), o);
      }
      cx.reduce<2>(s2);
      if (end_doubling) {
        nleap += st.n;
        accepted = accepted || swap;
        notdiv = st.nd;
        cont = cont_f && (s2[0] >= 0.f) && (s2[1] >= 0.f);
        to_chunk = cont && it + 1 < p.max_depth;
        it_next = it + 1; ihi_next = 0;
      } else if (live) {
        to_chunk = true;
        it_next = it; ihi_next = ihi + 1;
      }
    }
    if (sp.stats && tid == 0) atomicAdd(sp.stats + 2, (unsigned long long)(gt - gt_begin));
    lap(29);
    // ------------------------------------------------------------ (3) publish: finished transition or next task
    int next_cls = -1;   // -1: nothing to push (empty lane), 0 / 1: class, 2: all transitions done
    float gend[kK];
    cx.load_d(gend);   // gradient at the moving end (warp-aligned TMEM read, before the per-chain branches)
    if (live) {
      if (!to_chunk) {
        float fx[kK], fg[kK];
        rec_ld26(rv(kRCx), fx);
        rec_ld26(rv(kRCg), fg);
        const int leap = nleap * p.unrolled;
        const float lar = logf(esum / (float)nleap);
        const int r = tile_result_index(p, t);
        tile_store(p.x, 0, p.B, c, D, cx.slice, true, fx);
        tile_store(p.g, 0, p.B, c, D, cx.slice, true, fg);
        if (r >= 0) {
          const Trace& tr = p.tr;
          if (tr.states) tile_store(tr.states, r, p.B, c, D, cx.slice, true, fx);
          if (tr.grads) tile_store(tr.grads, r, p.B, c, D, cx.slice, true, fg);
        }
        if (cx.slice == 0) {
          p.lp[c] = clp;
          if (p.leapfrog_total) p.leapfrog_total[c] += (unsigned long long)leap;
          if (r >= 0) {
            const Trace& tr = p.tr;
            const size_t o = (size_t)r * p.B + c;
            if (tr.target_log_prob) tr.target_log_prob[o] = clp;
            if (tr.log_accept_ratio) tr.log_accept_ratio[o] = lar;
            if (tr.is_accepted) tr.is_accepted[o] = accepted ? 1 : 0;
            if (tr.leapfrogs_taken) tr.leapfrogs_taken[o] = leap;
            if (tr.has_divergence) tr.has_divergence[o] = notdiv ? 0 : 1;
            if (tr.reach_max_depth) tr.reach_max_depth[o] = cont ? 1 : 0;
            if (tr.energy) tr.energy[o] = cen;
            if (tr.step_size && c == 0 && p.step_kind == 0) tr.step_size[r] = p.step[0];
          }
          __stcg(&rs[kST], __int_as_float(t + 1));
        }
        next_cls = (t + 1 < p.t1) ? 0 : 2;
      } else {
        rec_st26(rv(kRSx), x); rec_st26(rv(kRSm), m); rec_st26(rv(kRSg), gend);
        if (ihi_next != 0) rec_st26(rv(kRRs), rho);
        if (cx.slice == 0) {
          __stcg(&rs[kSLp], lp); __stcg(&rs[kSH0], H0); __stcg(&rs[kSSlp], slp); __stcg(&rs[kSOlp], olp);
          __stcg(&rs[kSClp], clp); __stcg(&rs[kSCen], cen); __stcg(&rs[kSCw], cw); __stcg(&rs[kSEsum], esum);
          __stcg(&rs[kSNleap], __int_as_float(nleap));
          __stcg(&rs[kSFlags],
                 __int_as_float((notdiv ? 2 : 0) | (accepted ? 4 : 0) | (s_is_right ? 8 : 0) | (st.nd ? 16 : 0)));
          __stcg(&rs[kSIt], __int_as_float(it_next)); __stcg(&rs[kSIhi], __int_as_float(ihi_next));
          if (ihi_next != 0) {
            __stcg(&rs[kSBlp], st.blp); __stcg(&rs[kSBen], st.ben); __stcg(&rs[kSBw], st.bw);
            __stcg(&rs[kSEsub], st.esum_sub); __stcg(&rs[kSN], __int_as_float(st.n));
          }
        }
        next_cls = 1;
      }
    }
    __threadfence();
    __syncthreads();   // all four slices of every chain have written their part of the record
    if (live && cx.slice == 0) {
      if (next_cls == 2) {
        atomicAdd(sp.qctl + kQDone, 1ull);
      } else {
        const unsigned long long s = atomicAdd(sp.qctl + (next_cls ? kQTail1 : kQTail0), 1ull);
        *(volatile int*)(sp.queue + (size_t)next_cls * p.B + (size_t)(s % Bq)) = c;
      }
    }
    __syncthreads();
    if (tid == 0) atomicAdd(sp.qctl + kQRunning, (unsigned long long)(-(long long)ntask));
    lap(30);
  }
  cx.finish();
}

}  // namespace pb2
