// The persistent chain kernel's body, shared by the offline build (pb2_chain_kernels.cu: the named targets) and the
// run-time build of user-defined targets (pb2_user.cu compiles pb2_user_target.cuh with NVRTC): one thread group per
// chain, chains pulled from a dynamic queue, each group running transitions [t0,t1) of its chain back to back with the
// state in registers.
#pragma once
#include "pb2_chain.cuh"

namespace pb2 {

enum Mode : int { kModeLogpGrad = 0, kModeLeapfrog = 1, kModeHMC = 2, kModeNUTS = 3 };

// Extra pointers for the two primitive modes.
struct PrimIO {
  const float* m_in;
  const float* x_in;
  const float* lp_in;
  const float* g_in;
  float* m_out;
  float* x_out;
  float* lp_out;
  float* g_out;
  int L;
};

struct SmemPlan {
  int cta_floats;    // CTA-shared target data
  int red_floats;    // block-group reduction buffers
  int group_floats;  // per-group region (scratch + target + checkpoints)
  int tgt_off;       // offset of target group smem inside the group region
  int ck_off;        // offset of checkpoint store inside the group region (-1: global)
  int ck_floats;     // floats per checkpoint array (m or rho)
  int cold_off;      // block groups, NUTS: offset of the cold vectors inside the group region (-1: registers)
};

// The persistent chain loop of one thread group (the body of every chain kernel, offline- or run-time-compiled).
template <class Grp, int E, class Tgt, int MODE>
__device__ __forceinline__ void chain_body(const ChainParams& p, const typename Tgt::Params& tp, const PrimIO& io,
                                           const SmemPlan& plan) {
  extern __shared__ __align__(16) float smem[];
  float* cta = smem;
  Tgt tgt;
  tgt.init_cta(tp, cta);
  __syncthreads();
  float* gbase;
  if constexpr (Grp::kIsBlock) gbase = smem + plan.cta_floats + plan.red_floats;
  else gbase = smem + plan.cta_floats + (threadIdx.x / Grp::G) * plan.group_floats;
  auto make_group = [&]() {
    if constexpr (Grp::kIsBlock) return Grp(gbase, smem + plan.cta_floats);
    else return Grp(gbase);
  };
  Grp grp = make_group();
  tgt.init_group(tp, grp, cta, gbase + plan.tgt_off);
  Chain<Grp, E, Tgt> ch(grp, tgt, p);
  float* ckm = nullptr;
  float* ckr = nullptr;
  if constexpr (MODE == kModeNUTS) {
    if (plan.ck_off >= 0) {
      ckm = gbase + plan.ck_off;
    } else {
      ckm = p.ckpt_global + (size_t)blockIdx.x * 2 * plan.ck_floats;
    }
    ckr = ckm + plan.ck_floats;
  }
  while (true) {
    int cc = 0;
    if (grp.lane == 0) cc = atomicAdd(p.queue, 1);
    cc = grp.bcast_int(cc, 0);
    if (cc >= p.B) break;
    ch.c = cc;
    ch.cg = (uint64_t)p.chain_offset + (uint64_t)cc;
    float x[E], g[E], lp;
    if constexpr (MODE == kModeLogpGrad) {
      ch.load_vec(io.x_in, x);
      lp = tgt.logp_grad(grp, x, g);
      ch.store_vec(io.g_out, 0, g);
      if (grp.lane == 0) io.lp_out[cc] = lp;
    } else if constexpr (MODE == kModeLeapfrog) {
      float m[E], eps[E];
      ch.load_vec(io.m_in, m);
      ch.load_vec(io.x_in, x);
      ch.load_vec(io.g_in, g);
      lp = io.lp_in[cc];
      ch.load_eps(0, eps);
      ch.leapfrog(m, x, lp, g, eps, io.L);
      ch.store_vec(io.m_out, 0, m);
      ch.store_vec(io.x_out, 0, x);
      ch.store_vec(io.g_out, 0, g);
      if (grp.lane == 0) io.lp_out[cc] = lp;
    } else {
      ch.load_vec(p.x, x);
      ch.load_vec(p.g, g);
      lp = p.lp[cc];
      unsigned long long nleap_total = 0;
#pragma unroll 1
      for (int t = p.t0; t < p.t1; ++t) {
        const int r = ch.result_index(t);
        if constexpr (MODE == kModeHMC) {
          ch.hmc_transition(t, x, lp, g);
          nleap_total += (unsigned long long)p.L;
        } else {
          typename Chain<Grp, E, Tgt>::NutsOut no;
          ch.nuts_transition(t, x, lp, g, ckm, ckr, plan.cold_off >= 0 ? gbase + plan.cold_off : nullptr, no);
          nleap_total += (unsigned long long)no.leapfrogs;
          if (p.lar_last && grp.lane == 0) p.lar_last[cc] = no.log_accept_ratio;
          if (r >= 0 && grp.lane == 0) {
            const Trace& tr = p.tr;
            const size_t o = (size_t)r * p.B + cc;
            if (tr.log_accept_ratio) tr.log_accept_ratio[o] = no.log_accept_ratio;
            if (tr.is_accepted) tr.is_accepted[o] = no.accepted ? 1 : 0;
            if (tr.leapfrogs_taken) tr.leapfrogs_taken[o] = no.leapfrogs;
            if (tr.has_divergence) tr.has_divergence[o] = no.has_divergence ? 1 : 0;
            if (tr.reach_max_depth) tr.reach_max_depth[o] = no.reach_max_depth ? 1 : 0;
            if (tr.energy) tr.energy[o] = no.energy;
          }
        }
        if (r >= 0) {
          const Trace& tr = p.tr;
          if (tr.states) ch.store_vec(tr.states, r, x);
          if (tr.grads) ch.store_vec(tr.grads, r, g);
          if (grp.lane == 0) {
            if (tr.target_log_prob) tr.target_log_prob[(size_t)r * p.B + cc] = lp;
            if (tr.step_size && cc == 0 && p.step_kind == 0)
              tr.step_size[r] = p.step[(size_t)t * p.step_seq_stride];
          }
        }
      }
      ch.store_vec(p.x, 0, x);
      ch.store_vec(p.g, 0, g);
      if (grp.lane == 0) {
        p.lp[cc] = lp;
        if (p.leapfrog_total) p.leapfrog_total[cc] += nleap_total;
      }
    }
  }
}


}  // namespace pb2
