// Per-chain transition code: leapfrog, HMC (+ Metropolis-Hastings) and the
// iterative (unrolled) NUTS tree, executed by one thread group per chain with the
// chain state resident in registers for the whole transition (all L leapfrogs /
// all 2^depth tree steps).  Follows, op for op,
//   tfp/mcmc/internal/leapfrog_integrator.py:280-309,330-355   (leapfrog)
//   tfp/mcmc/hmc.py:661-729,780-875; metropolis_hastings.py:181-254; util.py:205-235
//   tfp/mcmc/nuts.py:321-445,512-946,949-1010                  (NUTS)
// Per-chain early exit is result-identical to the reference's lock-step masking
// because every key is a function of the step seed only.
#pragma once
#include "pb2_compat.cuh"
#include "pb2_group.cuh"
#include "pb2_rng.cuh"

namespace pb2 {

constexpr int kMaxParts = 8;

struct Trace {
  float* states;                 // [R,B,D]
  float* target_log_prob;        // [R,B]
  float* grads;                  // [R,B,D]
  float* log_accept_ratio;       // [R,B]
  unsigned char* is_accepted;    // [R,B]
  float* step_size;              // [R]  (scalar step size only)
  // HMC proposal details (MetropolisHastingsKernelResults.proposed_*)
  float* proposed_state;             // [R,B,D]
  float* proposed_target_log_prob;   // [R,B]
  float* proposed_grads;             // [R,B,D]
  float* log_acceptance_correction;  // [R,B]
  float* initial_momentum;           // [R,B,D]
  float* final_momentum;             // [R,B,D]
  // NUTS
  int* leapfrogs_taken;          // [R,B]
  unsigned char* has_divergence; // [R,B]
  unsigned char* reach_max_depth;// [R,B]
  float* energy;                 // [R,B]
};

struct ChainParams {
  int B, D, B_global, chain_offset, layout;
  int n_parts;
  int part_off[kMaxParts + 1];
  float* x;   // [B,D] in/out
  float* lp;  // [B]   in/out
  float* g;   // [B,D] in/out
  const float* step;    // scalar: step[t*step_seq_stride]; per-dim: [D]; per-chain: [B]
  int step_kind;        // 0 scalar, 1 per-dim, 2 per-chain
  int step_seq_stride;
  int t0, t1;           // transitions [t0,t1) are run by this launch
  int t_sched0;         // transition index of schedule row 0
  float* lar_last;      // [B] optional: log_accept_ratio of the latest transition (dual averaging)
  int burnin, thin, n_results;
  int* queue;                           // dynamic chain queue (zeroed before launch)
  unsigned long long* leapfrog_total;   // [B] optional, += leapfrogs of every transition
  int L;                                // HMC leapfrogs per transition
  int max_depth;                        // NUTS
  float max_energy_diff;
  int unrolled;
  const uint32_t* sched;  // per-transition key schedule (see pb2_sched kernels)
  int sched_stride;       // uint32 words per transition
  float* ckpt_global;     // block groups: [gridDim.x][2*max_depth*E*G]
  const float* scale;     // (nullable, [D]) diagonal preconditioning: the kernels run on u = x / scale (pb2_targets.cuh ScaledT)
  const int* bij_kind;    // (nullable, [D]) per-dimension event-space bijectors: the chains live in the unconstrained
  const float* bij_lo;    //   space of a TransformedTransitionKernel (pb2_targets.cuh TransformedT)
  const float* bij_hi;
  Trace tr;
};

__device__ __forceinline__ float finite_or_neginf(float v) { return isfinite(v) ? v : -INFINITY; }

__device__ __forceinline__ float log_add_exp(float x, float y) {
  // tfp/math/generic.py:585-611
  float larger = fmaxf(x, y);
  float t = (x - larger) + (y - larger);
  return larger + (log1pf(expf(-fabsf(t))) + fmaxf(t, 0.f));
}

template <class Grp, int E, class Tgt>
struct Chain {
  Grp& grp;
  Tgt& tgt;
  const ChainParams& p;
  int c;         // local chain
  uint64_t cg;   // global chain index (RNG counter row)

  __device__ Chain(Grp& g_, Tgt& t_, const ChainParams& p_) : grp(g_), tgt(t_), p(p_), c(0), cg(0) {}

  __device__ __forceinline__ int dim(int j) const { return grp.lane * E + j; }

  __device__ __forceinline__ void load_vec(const float* base, float (&v)[E]) const {
    const float* row = base + (size_t)c * p.D;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      int d = dim(j);
      v[j] = d < p.D ? row[d] : 0.f;
    }
  }
  __device__ __forceinline__ void store_vec(float* base, size_t r, const float (&v)[E]) const {
    float* row = base + (r * (size_t)p.B + c) * p.D;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      int d = dim(j);
      if (d < p.D) row[d] = v[j];
    }
  }
  __device__ __forceinline__ void load_eps(int t, float (&eps)[E]) const {
#pragma unroll
    for (int j = 0; j < E; ++j) {
      int d = dim(j);
      float e;
      if (p.step_kind == 0) e = p.step[(size_t)t * p.step_seq_stride];
      else if (p.step_kind == 1) e = d < p.D ? p.step[d] : 0.f;
      else e = p.step[c];
      eps[j] = e;
    }
  }

  // momentum ~ N(0, I): one key per state part, counter = row-major index in [B_global, size_p]
  __device__ __forceinline__ void draw_momentum(const uint32_t* keys, float (&m)[E]) const {
#pragma unroll
    for (int j = 0; j < E; ++j) {
      int d = dim(j);
      float val = 0.f;
      if (d < p.D) {
        int part = 0;
#pragma unroll 1
        for (int q = 1; q < p.n_parts; ++q) part += (d >= p.part_off[q]) ? 1 : 0;
        const int off = p.part_off[part];
        const uint64_t sz = (uint64_t)(p.part_off[part + 1] - off);
        Key k{keys[2 * part], keys[2 * part + 1]};
        uint32_t b = bits_at(k, cg * sz + (uint64_t)(d - off), (uint64_t)p.B_global * sz, p.layout);
        val = normal_from_bits(b);
      }
      m[j] = val;
    }
  }

  __device__ __forceinline__ uint32_t chain_bits(const uint32_t* key2) const {
    Key k{key2[0], key2[1]};
    return bits_at(k, cg, (uint64_t)p.B_global, p.layout);
  }

  // leapfrog_integrator.py:280-309: v = m + (0.5*eps) g; L x {x += eps v; (lp,g)=f(x); v += eps g}; m = v - (0.5*eps) g
  __device__ __forceinline__ void leapfrog(float (&m)[E], float (&x)[E], float& lp, float (&g)[E],
                                           const float (&eps)[E], int L) {
    float v[E];
#pragma unroll
    for (int j = 0; j < E; ++j) v[j] = m[j] + (0.5f * eps[j]) * g[j];
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
#pragma unroll
      for (int j = 0; j < E; ++j) x[j] = x[j] + eps[j] * v[j];
      lp = tgt.logp_grad(grp, x, g);
#pragma unroll
      for (int j = 0; j < E; ++j) v[j] = v[j] + eps[j] * g[j];
    }
#pragma unroll
    for (int j = 0; j < E; ++j) m[j] = v[j] - (0.5f * eps[j]) * g[j];
  }

  __device__ __forceinline__ float sumsq(const float (&m)[E]) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) s = fmaf(m[j], m[j], s);
    return grp.sum(s);
  }

  __device__ __forceinline__ int result_index(int t) const {
    // sample.py:359-366: first result after 1+burnin steps, then every 1+thin
    int u = t - p.burnin;
    if (u < 0) return -1;
    int q = u / (p.thin + 1);
    if (q * (p.thin + 1) != u || q >= p.n_results) return -1;
    return q;
  }

  // ---------------------------------------------------------------- HMC + MH
  __device__ void hmc_transition(int t, float (&x)[E], float& lp, float (&g)[E]) {
    const uint32_t* sk = p.sched + (size_t)(t - p.t_sched0) * p.sched_stride;
    float eps[E];
    load_eps(t, eps);
    float m0[E], m1[E], x1[E], g1[E];
    draw_momentum(sk, m0);
#pragma unroll
    for (int j = 0; j < E; ++j) { m1[j] = m0[j]; x1[j] = x[j]; g1[j] = g[j]; }
    float lp1 = lp;
    leapfrog(m1, x1, lp1, g1, eps, p.L);
    float ks[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < E; ++j) { ks[0] = fmaf(m0[j], m0[j], ks[0]); ks[1] = fmaf(m1[j], m1[j], ks[1]); }
    grp.template sumN<2>(ks);
    const float corr = 0.5f * finite_or_neginf(ks[0] + (-ks[1]));      // hmc.py:862-875
    const float ratio = finite_or_neginf((lp1 + (-lp)) + corr);       // metropolis_hastings.py:204-215
    const float u = uniform_from_bits(chain_bits(sk + 2 * p.n_parts), 0.f, 1.f);
    const bool accept = logf(u) < ratio;                               // :221-227
    const int r = result_index(t);
    if (p.lar_last && grp.lane == 0) p.lar_last[c] = ratio;
    if (r >= 0) {
      const Trace& tr = p.tr;
      if (tr.proposed_state) store_vec(tr.proposed_state, r, x1);
      if (tr.proposed_grads) store_vec(tr.proposed_grads, r, g1);
      if (tr.initial_momentum) store_vec(tr.initial_momentum, r, m0);
      if (tr.final_momentum) store_vec(tr.final_momentum, r, m1);
      if (grp.lane == 0) {
        size_t o = (size_t)r * p.B + c;
        if (tr.proposed_target_log_prob) tr.proposed_target_log_prob[o] = lp1;
        if (tr.log_acceptance_correction) tr.log_acceptance_correction[o] = corr;
        if (tr.log_accept_ratio) tr.log_accept_ratio[o] = ratio;
        if (tr.is_accepted) tr.is_accepted[o] = accept ? 1 : 0;
      }
    }
    if (accept) {
#pragma unroll
      for (int j = 0; j < E; ++j) { x[j] = x1[j]; g[j] = g1[j]; }
      lp = lp1;
    }
  }

  // ---------------------------------------------------------------- NUTS
  // checkpoint store: [slot][j][lane]  (conflict-free in smem, coalesced in global)
  __device__ __forceinline__ static int ck_idx(int slot, int j, int lane) { return (slot * E + j) * Grp::G + lane; }

  static constexpr int kChunk = Grp::G < 32 ? Grp::G : 32;

  struct NutsOut {
    float energy, log_accept_ratio;
    int leapfrogs;
    bool accepted, reach_max_depth, has_divergence;
  };

  static constexpr int kColdVectors = 7;   // om, ox, og, cx, cgd, bx, bg
  __device__ void nuts_transition(int t, float (&x)[E], float& lp, float (&g)[E], float* ckm, float* ckr,
                                  float* cold, NutsOut& out) {
    const uint32_t* sk = p.sched + (size_t)(t - p.t_sched0) * p.sched_stride;
    const uint32_t* hdr = sk + 2 * p.n_parts;                  // per depth: dir key, acc key, sub key
    const uint32_t* ku = hdr + 6 * p.max_depth;                // subtree uniform keys, depth j at 2^j - 1
    float* rb = grp.scratch;                                   // [0,16) dir, [16,32) log1p(-u_acc), [32,64) subtree log1p(-u)
    float eps_abs[E];
    load_eps(t, eps_abs);
    // per-depth randoms for this chain, lane-parallel (nuts.py:551-558, :622-625)
    grp.sync();
    if (grp.lane < p.max_depth) {
      const int j = grp.lane;
      rb[j] = (float)(chain_bits(hdr + 6 * j) & 1u);
      rb[16 + j] = log1pf(-uniform_from_bits(chain_bits(hdr + 6 * j + 2), 0.f, 1.f));
    }
    using ColdV = typename Grp::template Cold<E>;
    float sm[E], sx[E], sg[E], slp;      // moving end
    ColdV om(cold, 0, grp.lane), ox(cold, 1, grp.lane), og(cold, 2, grp.lane);   // other end
    float olp;
    ColdV cx(cold, 3, grp.lane), cgd(cold, 4, grp.lane);                         // candidate
    float clp, cen, cw;
    ColdV bx(cold, 5, grp.lane), bg(cold, 6, grp.lane);                          // subtree candidate
    float blp, ben, bw;
    float rho[E], rhos[E];
    draw_momentum(sk, sm);               // nuts.py:515-523
    const float H0 = lp - 0.5f * sumsq(sm);  // compute_hamiltonian :1085-1102
#pragma unroll
    for (int j = 0; j < E; ++j) {
      sx[j] = x[j]; sg[j] = g[j];
      om[j] = sm[j]; ox[j] = x[j]; og[j] = g[j];
      cx[j] = x[j]; cgd[j] = g[j];
      rho[j] = sm[j];
    }
    slp = lp; olp = lp; clp = lp; cen = H0; cw = 0.f;
    bool s_is_right = true;  // which trajectory end the registers `s*` currently hold
    float esum = 0.f;
    int nleap = 0;
    bool cont = true, notdiv = true, accepted = false;
    grp.sync();
#pragma unroll 1
    for (int it = 0; it < p.max_depth && cont; ++it) {
      const bool dir = rb[it] != 0.f;  // true: extend the right end with +eps (nuts.py:560-574)
      if (dir != s_is_right) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
          float a;
          a = sm[j]; sm[j] = om[j]; om[j] = a;
          a = sx[j]; sx[j] = ox[j]; ox[j] = a;
          a = sg[j]; sg[j] = og[j]; og[j] = a;
        }
        float a = slp; slp = olp; olp = a;
        s_is_right = dir;
      }
      float eps[E];
#pragma unroll
      for (int j = 0; j < E; ++j) eps[j] = dir ? eps_abs[j] : -eps_abs[j];
      // _build_sub_tree nuts.py:713-791
#pragma unroll
      for (int j = 0; j < E; ++j) { bx[j] = sx[j]; bg[j] = sg[j]; rhos[j] = 0.f; }
      blp = slp; ben = slp; bw = -INFINITY;
      int n = 0;
      bool c_prev = true, nd = notdiv;
      float esum_sub = 0.f;
      const int nsteps = 1 << it;
      const uint32_t* kud = ku + 2 * (nsteps - 1);
#pragma unroll 1
      for (int i = 0; i < nsteps && c_prev; ++i) {
        if ((i & (kChunk - 1)) == 0) {  // next kChunk multinomial uniforms of this chain, lane-parallel
          grp.sync();
          if (grp.lane < kChunk && i + grp.lane < nsteps)
            rb[32 + grp.lane] = log1pf(-uniform_from_bits(chain_bits(kud + 2 * (i + grp.lane)), 0.f, 1.f));
          grp.sync();
        }
        leapfrog(sm, sx, slp, sg, eps, p.unrolled);  // _loop_build_sub_tree :816-819
        n += 1;
        bool ok = true;
        float ksum;
        if ((i & 1) == 0) {
          // even step: store (m', rho_prev) at slot popcount(i); no checks (nuts.py:858-869, N1 closed form)
          const int slot = __popc(i);
          float kk = 0.f;
#pragma unroll
          for (int j = 0; j < E; ++j) {
            ckm[ck_idx(slot, j, grp.lane)] = sm[j];
            ckr[ck_idx(slot, j, grp.lane)] = rhos[j];
            rhos[j] += sm[j];
            kk = fmaf(sm[j], sm[j], kk);
          }
          ksum = grp.sum(kk);
        } else {
          // odd step: U-turn checks against slots [pc - trailing_ones, pc) (nuts.py:949-1010)
#pragma unroll
          for (int j = 0; j < E; ++j) rhos[j] += sm[j];
          const int pc = __popc(i);
          const int k0 = pc - (__ffs(~i) - 1);
          float kk = 0.f;
#pragma unroll
          for (int j = 0; j < E; ++j) kk = fmaf(sm[j], sm[j], kk);
          float s3[3] = {kk, 0.f, 0.f};
          {
#pragma unroll
            for (int j = 0; j < E; ++j) {
              const float diff = rhos[j] - ckr[ck_idx(k0, j, grp.lane)];
              s3[1] = fmaf(diff, ckm[ck_idx(k0, j, grp.lane)], s3[1]);
              s3[2] = fmaf(diff, sm[j], s3[2]);
            }
          }
          grp.template sumN<3>(s3);
          ksum = s3[0];
          ok = (s3[1] >= 0.f) && (s3[2] >= 0.f);
#pragma unroll 1
          for (int k = k0 + 1; k < pc && ok; ++k) {
            float s2[2] = {0.f, 0.f};
#pragma unroll
            for (int j = 0; j < E; ++j) {
              const float diff = rhos[j] - ckr[ck_idx(k, j, grp.lane)];
              s2[0] = fmaf(diff, ckm[ck_idx(k, j, grp.lane)], s2[0]);
              s2[1] = fmaf(diff, sm[j], s2[1]);
            }
            grp.template sumN<2>(s2);
            ok = (s2[0] >= 0.f) && (s2[1] >= 0.f);
          }
        }
        float en = slp - 0.5f * ksum;                 // :871-877
        en = isnan(en) ? -INFINITY : en;
        const float dH = en - H0;
        const bool nd_i = (-dH) < p.max_energy_diff;  // :880
        const float w_new = log_add_exp(bw, dH);      // :881-883
        const bool take = rb[32 + (i & (kChunk - 1))] <= (dH - w_new);  // :897-901
        if (take) {
#pragma unroll
          for (int j = 0; j < E; ++j) { bx[j] = sx[j]; bg[j] = sg[j]; }
          blp = slp; ben = en;
        }
        bw = w_new;
        const bool c_now = nd_i;                      // & c_prev (true inside the loop) :921
        if (c_now) esum_sub += expf(fminf(dH, 0.f));  // :930-933
        nd = nd && nd_i;                              // :924-927,944
        c_prev = ok && c_now;                         // :922
      }
      const bool cont_f = c_prev;
      // _loop_tree_doubling nuts.py:597-711
      esum = esum_sub + esum;
      const float tw = cont_f ? bw : -INFINITY;
      const float wsum = log_add_exp(tw, cw);
      float thr = tw - cw;
      thr = isnan(thr) ? 0.f : thr;
      const bool swap = (rb[16 + it] <= thr) && cont_f;
      cw = wsum;
      if (swap) {
#pragma unroll
        for (int j = 0; j < E; ++j) { cx[j] = bx[j]; cgd[j] = bg[j]; }
        clp = blp; cen = ben;
      }
      float s2[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < E; ++j) {
        rho[j] += rhos[j];
        s2[0] = fmaf(rho[j], sm[j], s2[0]);
        s2[1] = fmaf(rho[j], om[j], s2[1]);
      }
      grp.template sumN<2>(s2);
      nleap += n;
      accepted = accepted || swap;
      notdiv = nd;
      cont = cont_f && (s2[0] >= 0.f) && (s2[1] >= 0.f);
    }
#pragma unroll
    for (int j = 0; j < E; ++j) { x[j] = cx[j]; g[j] = cgd[j]; }
    lp = clp;
    out.energy = cen;
    out.log_accept_ratio = logf(esum / (float)nleap);   // nuts.py:429-432
    out.leapfrogs = nleap * p.unrolled;
    out.accepted = accepted;
    out.reach_max_depth = cont;
    out.has_divergence = !notdiv;
  }
};

}  // namespace pb2
