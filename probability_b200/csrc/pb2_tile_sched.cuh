// tile_nuts_sched_kernel: the tcgen05 tile NUTS kernel with chains re-grouped at DOUBLING boundaries.
//
// The lock-step tile kernel (tile_nuts_kernel) makes all 128 chains of a tile wait for the tile's
// deepest tree (measured utilisation 0.27 on the 100-d ill-conditioned Gaussian: mean 276 of max
// 1023 leapfrogs).  Here the unit of work is a TASK = one tree doubling (2^it leaves, identical for
// every chain of the task):
//   class 0      : start a transition (momentum, H0, ...) and run doublings it = 0 .. s0-1
//   class k >= 1 : doubling it = s0 + k - 1
// Persistent CTAs repeatedly (1) histogram the per-chain `ready` words, (2) claim up to 128 chains of
// one class with atomicCAS, (3) run the task in lock-step exactly like tile_nuts_kernel, (4) publish
// each chain's next class (or finish its transition, emit the traced results and start the next one).
// A chain's results do not depend on which tile/CTA ran its tasks: every random number is a function
// of (step seed, global chain index) and all per-chain reductions have a fixed order.
// Chain-major records hold what survives a doubling boundary (both trajectory ends, the trajectory
// candidate, rho and the scalars); the subtree candidate and the checkpoint stores are dead at a
// boundary and stay in the CTA's scratch.
#pragma once
#include "pb2_tile_nuts.cuh"

namespace pb2 {
using namespace tile;

enum { kRSx = 0, kRSm, kRSg, kROx, kROm, kROg, kRCx, kRCg, kRRho, kRecVecs };
enum { kSLp = 0, kSH0, kSSlp, kSOlp, kSClp, kSCen, kSCw, kSEsum, kSNleap, kSFlags, kST, kSNleapTot, kRecScal = 16 };
constexpr int kReadyRunning = -1, kReadyDone = 99;

struct SchedParams {
  float* rec_vec;    // [B][kRecVecs][kKP]
  float* rec_scal;   // [B][kRecScal]
  int* ready;        // [B]
  int s0;            // doublings merged into the start task
  int patience;      // polls without a full tile before a partial tile is accepted
  unsigned long long* stats;   // optional [32]: tasks, claimed chains, ticks, idle polls, per-class tasks @8+
};

__global__ void tile_sched_init_kernel(int* ready, float* rec_scal, int B, int t0) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B) return;
  ready[c] = 0;
  rec_scal[(size_t)c * kRecScal + kST] = __int_as_float(t0);
  rec_scal[(size_t)c * kRecScal + kSNleapTot] = __int_as_float(0);
}

__global__ void __launch_bounds__(kThreads, 1)
tile_nuts_sched_kernel(const ChainParams p, const DenseGaussianParams tp, const SchedParams sp,
                       float* __restrict__ scratch_all) {
  extern __shared__ __align__(128) unsigned char planes[];
  __shared__ Shared sh;
  __shared__ float lu[4][kM];
  __shared__ int hist[16];
  __shared__ int ids[kM];
  __shared__ int n_claimed, n_done, sel_cls;
  Ctx cx;
  cx.init(&sh, planes, tp.P, tp.loc, tp.D);
  Prof pf;
  pf.init();
  const int D = tp.D;
  const int tid = threadIdx.x;
  const int cl = cx.cl;
  // CTA scratch: subtree candidate (x, g) + checkpoint stores; per-thread segments (pb2_tile.cuh seg_*)
  enum { kTBx = 0, kTBg = 1, kTCk = 2 };
  const int nvec = kTCk + 2 * p.max_depth;
  float* const scr_s = scratch_all + (size_t)blockIdx.x * nvec * kVS + (size_t)(kK * cx.slice) * kM;
  SubtreeArgs sa;
  sa.unrolled = p.unrolled; sa.layout = p.layout; sa.b_global = (uint64_t)p.B_global;
  sa.lognorm = tp.lognorm; sa.max_energy_diff = p.max_energy_diff;
  sa.lc = sh.loc + kK * cx.slice;
  sa.bx = scr_s + (size_t)kTBx * kVS; sa.bg = scr_s + (size_t)kTBg * kVS; sa.ck = scr_s + (size_t)kTCk * kVS;
  sa.max_depth = p.max_depth;
  sa.ckl = reinterpret_cast<float*>(planes + 2 * kPlaneBytes) + (size_t)(kK * cx.slice) * kM;
  const int nclass = 1 + max(0, p.max_depth - sp.s0);
  unsigned gt = 0;
  int patience = 0;                 // consecutive polls without a full tile (uniform across the CTA)
  const int scan0 = (int)(((long long)blockIdx.x * p.B) / gridDim.x);   // de-correlate the CTAs' scans

  long long tmark = clock64();
  auto lap = [&](int k) {   // thread 0 of every CTA: cycles per scheduler phase (only with PB2_SCHED_STATS)
    if (sp.stats && tid == 0) {
      const long long now = clock64();
      atomicAdd(sp.stats + k, (unsigned long long)(now - tmark));
      tmark = now;
    }
  };
  while (true) {
    // ------------------------------------------------------------ (1) what is ready?
    __syncthreads();   // everybody is done reading the shared scheduling words of the previous round
    if (tid < 16) hist[tid] = 0;
    if (tid == 0) { n_claimed = 0; n_done = 0; }
    __syncthreads();
    {
      int loc_done = 0;
      for (int i = tid; i < p.B; i += kThreads) {
        const int rd = *(volatile int*)(sp.ready + i);
        if (rd == kReadyDone) loc_done++;
        else if (rd >= 0 && rd < 16) atomicAdd(&hist[rd], 1);
      }
      if (loc_done) atomicAdd(&n_done, loc_done);
    }
    __syncthreads();
    if (n_done == p.B) break;
    if (tid == 0) {
      // policy: a FULL tile of the deepest class first (deep tasks are long: running them partially filled is
      // what wastes the machine).  Without a full tile: wait for producers for a while (patience), then take
      // the fullest class -- immediately if nothing is running anywhere (end of the run).
      int best = -1, best_n = 0, n_ready = 0;
      for (int k = 0; k < nclass; ++k) n_ready += hist[k];
      for (int k = nclass - 1; k >= 0; --k)
        if (hist[k] >= kM) { best = k; best_n = hist[k]; break; }
      if (best < 0) {
        const int n_running = p.B - n_done - n_ready;
        if (n_running == 0 || patience >= sp.patience) {
          for (int k = 0; k < nclass; ++k)
            if (hist[k] > best_n) { best = k; best_n = hist[k]; }
        }
      }
      sel_cls = best;
    }
    __syncthreads();
    const int cls = sel_cls;
    if (cls < 0) {
      if (sp.stats && tid == 0) atomicAdd(sp.stats + 3, 1ull);
      patience++;
      __nanosleep(10000);
      lap(27);
      continue;
    }
    // ------------------------------------------------------------ (2) claim up to 128 chains of that class
    for (int i0 = tid; i0 < p.B; i0 += kThreads) {
      if (*(volatile int*)&n_claimed >= kM) break;
      int i = i0 + scan0;
      if (i >= p.B) i -= p.B;
      if (*(volatile int*)(sp.ready + i) == cls && atomicCAS(sp.ready + i, cls, kReadyRunning) == cls) {
        const int slot = atomicAdd(&n_claimed, 1);
        if (slot < kM) ids[slot] = i;
        else atomicExch(sp.ready + i, cls);          // tile is full: give it back
      }
    }
    __syncthreads();
    const int ntask = min(n_claimed, kM);
    if (ntask == 0) { lap(27); continue; }
    lap(26);
    patience = 0;
    if (sp.stats && tid == 0) {
      atomicAdd(sp.stats + 0, 1ull);
      atomicAdd(sp.stats + 1, (unsigned long long)ntask);
      atomicAdd(sp.stats + 8 + cls, 1ull);
      atomicAdd(sp.stats + 20 + cls, (unsigned long long)ntask);
    }
    const unsigned gt_begin = gt;
    __threadfence();
    // ------------------------------------------------------------ (3) run the task in lock-step
    const bool live = cx.cl < ntask;
    const int c = live ? ids[cx.cl] : 0;
    const uint64_t cg = (uint64_t)p.chain_offset + (uint64_t)c;
    float* const rec = sp.rec_vec + (size_t)c * kRecVecs * kKP + kK * cx.slice;   // element j of vector v: rec[v*kKP + j]
    float* const rs = sp.rec_scal + (size_t)c * kRecScal;
    const int t = live ? __float_as_int(__ldcg(&rs[kST])) : p.t0;
    const float eps_abs = p.step_kind == 0 ? p.step[0] : (live ? p.step[c] : 0.f);
    const uint32_t* sk = p.sched + (size_t)(t - p.t_sched0) * p.sched_stride;
    const uint32_t* hdr = sk + 2 * p.n_parts;
    const uint32_t* ku = hdr + 6 * p.max_depth;
    sa.cg = cg;
    float x[kK], m[kK], rho[kK];
    float lp, H0, slp, olp, clp, cen, cw, esum;
    int nleap;
    bool cont, notdiv, accepted, s_is_right;
    int it_begin, it_end;
    if (cls == 0) {
      // ---- _start_trajectory_batched (nuts.py:512-539)
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
        if (live) {
          rec[kROx * kKP + j] = x[j]; rec[kROm * kKP + j] = mm; rec[kROg * kKP + j] = g[j];
          rec[kRCx * kKP + j] = x[j]; rec[kRCg * kKP + j] = g[j];
          rec[kRRho * kKP + j] = mm;
        }
      }
      cx.reduce<1>(s1);
      H0 = lp - 0.5f * s1[0];
      slp = lp; olp = lp; clp = lp; cen = H0; cw = 0.f; esum = 0.f;
      nleap = 0;
      cont = live; notdiv = true; accepted = false; s_is_right = true;
      it_begin = 0;
      it_end = min(sp.s0, p.max_depth);
    } else {
      float g[kK];
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        x[j] = live ? __ldcg(&rec[kRSx * kKP + j]) : 0.f;
        m[j] = live ? __ldcg(&rec[kRSm * kKP + j]) : 0.f;
        g[j] = live ? __ldcg(&rec[kRSg * kKP + j]) : 0.f;
      }
      cx.store_d(g);
      lp = live ? __ldcg(&rs[kSLp]) : 0.f;  H0 = live ? __ldcg(&rs[kSH0]) : 0.f;  slp = live ? __ldcg(&rs[kSSlp]) : 0.f;
      olp = live ? __ldcg(&rs[kSOlp]) : 0.f;  clp = live ? __ldcg(&rs[kSClp]) : 0.f;  cen = live ? __ldcg(&rs[kSCen]) : 0.f;
      cw = live ? __ldcg(&rs[kSCw]) : 0.f;  esum = live ? __ldcg(&rs[kSEsum]) : 0.f;
      nleap = live ? __float_as_int(rs[kSNleap]) : 0;
      const int fl = live ? __float_as_int(rs[kSFlags]) : 0;
      cont = live; notdiv = (fl & 2) != 0; accepted = (fl & 4) != 0; s_is_right = (fl & 8) != 0;
      it_begin = sp.s0 + cls - 1;
      it_end = it_begin + 1;
    }
    sa.H0 = H0;
    int any_cont = __syncthreads_or(cont ? 1 : 0);
    lap(28);
#pragma unroll 1
    for (int it = it_begin; it < it_end && any_cont; ++it) {
      Key kd{hdr[6 * it], hdr[6 * it + 1]}, kac{hdr[6 * it + 2], hdr[6 * it + 3]};
      const bool dir = (bits_at(kd, cg, (uint64_t)p.B_global, p.layout) & 1u) != 0;
      const float lacc = log1pf(-uniform_from_bits(bits_at(kac, cg, (uint64_t)p.B_global, p.layout), 0.f, 1.f));
      {
        const bool sw = live && dir != s_is_right;
        float g[kK];
        cx.load_d(g);
        if (sw) {
#pragma unroll
          for (int j = 0; j < kK; ++j) {
            float a;
            a = __ldcg(&rec[kROx * kKP + j]); rec[kROx * kKP + j] = x[j]; x[j] = a;
            a = __ldcg(&rec[kROm * kKP + j]); rec[kROm * kKP + j] = m[j]; m[j] = a;
            a = __ldcg(&rec[kROg * kKP + j]); rec[kROg * kKP + j] = g[j]; g[j] = a;
          }
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
      SubtreeState st;
      st.slp = slp; st.c_prev = cont; st.nd = notdiv;
      nuts_subtree(cx, sh, lu, gt, sa, x, m, rho, st, pf);
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
        seg_ld26(sa.bx, cl, o);
#pragma unroll
        for (int j = 0; j < kK; ++j) rec[kRCx * kKP + j] = o[j];
        seg_ld26(sa.bg, cl, o);
#pragma unroll
        for (int j = 0; j < kK; ++j) rec[kRCg * kKP + j] = o[j];
        clp = st.blp; cen = st.ben;
      }
      float s2[2] = {0.f, 0.f};
      if (live) {
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          const float rr = __ldcg(&rec[kRRho * kKP + j]) + rho[j];
          rec[kRRho * kKP + j] = rr;
          s2[0] = fmaf(rr, m[j], s2[0]);
          s2[1] = fmaf(rr, __ldcg(&rec[kROm * kKP + j]), s2[1]);
        }
      }
      cx.reduce<2>(s2);
      nleap += st.n;
      accepted = accepted || swap;
      notdiv = st.nd;
      cont = cont_f && (s2[0] >= 0.f) && (s2[1] >= 0.f);
      any_cont = __syncthreads_or(cont ? 1 : 0);
    }
    if (sp.stats && tid == 0) atomicAdd(sp.stats + 2, (unsigned long long)(gt - gt_begin));
    lap(29);
    // ------------------------------------------------------------ (4) publish: finished transition or next doubling
    const bool finished = !cont || it_end >= p.max_depth;
    int next_ready = kReadyRunning;
    float gend[kK];
    cx.load_d(gend);   // gradient at the moving end (warp-aligned TMEM read, before the per-chain branches)
    if (live) {
      if (finished) {
        float fx[kK], fg[kK];
#pragma unroll
        for (int j = 0; j < kK; ++j) { fx[j] = __ldcg(&rec[kRCx * kKP + j]); fg[j] = __ldcg(&rec[kRCg * kKP + j]); }
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
          rs[kST] = __int_as_float(t + 1);
        }
        next_ready = (t + 1 < p.t1) ? 0 : kReadyDone;
      } else {
#pragma unroll
        for (int j = 0; j < kK; ++j) {
          rec[kRSx * kKP + j] = x[j]; rec[kRSm * kKP + j] = m[j]; rec[kRSg * kKP + j] = gend[j];
        }
        if (cx.slice == 0) {
          rs[kSLp] = lp; rs[kSH0] = H0; rs[kSSlp] = slp; rs[kSOlp] = olp; rs[kSClp] = clp; rs[kSCen] = cen;
          rs[kSCw] = cw; rs[kSEsum] = esum;
          rs[kSNleap] = __int_as_float(nleap);
          rs[kSFlags] = __int_as_float((notdiv ? 2 : 0) | (accepted ? 4 : 0) | (s_is_right ? 8 : 0));
        }
        next_ready = (cls == 0) ? 1 : cls + 1;
      }
    }
    __threadfence();
    __syncthreads();   // all four slices of every chain have written their part of the record
    if (live && cx.slice == 0) atomicExch(sp.ready + c, next_ready);
    lap(30);
  }
  cx.finish();
}

}  // namespace pb2
