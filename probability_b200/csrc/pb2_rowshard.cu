// Lock-step (all chains at once) kernels for the large-data, ROW-SHARDED logistic regression
// (BASELINE config 5: N = 1e6 rows split over the ranks, 1,024 replicated chains, the per-leapfrog
// gradient all-reduced over NVLink by the caller).  Reference semantics: a target whose
// log-likelihood is psum'd over the data axis (tfp/internal/distribute_lib.py:179-242) evaluated
// inside SimpleLeapfrogIntegrator (leapfrog_integrator.py:280-355).
//
// rowshard_logistic_kernel: thread = chain, theta[DP] and the gradient accumulator g[DP] live in
// registers for the whole pass; a CTA of 128 chains streams its slice of X through shared memory
// (cp.async double buffer); every row is read from smem as broadcast LDS.128.  Per (row, chain):
// 2*D FMA + 1 exp + 1 log + 1 rcp.  Partials per row-group are summed by a second, deterministic
// kernel (fixed order => identical bits on every launch; after the all-reduce every rank holds the
// same gradient, so replicated chains take identical accept decisions).
#include <cuda_pipeline.h>
#include "pb2_internal.h"

namespace pb2 {

constexpr int kRsThreads = 128;  // chains per CTA
constexpr int kRsRows = 32;      // rows per smem stage

template <int DP>
__global__ void __launch_bounds__(kRsThreads, 2)
rowshard_logistic_kernel(const float* __restrict__ X, const float* __restrict__ y, int N, int D,
                         const float* __restrict__ theta, int B, int rows_per_group,
                         float* __restrict__ part_g /*[RG,B,D]*/, float* __restrict__ part_ll /*[RG,B]*/) {
  __shared__ __align__(16) float xs[2][kRsRows * DP];
  __shared__ float ys[2][kRsRows];
  const int b = blockIdx.x * kRsThreads + threadIdx.x;
  const int rg = blockIdx.y;
  const int n0 = rg * rows_per_group;
  const int n1 = min(N, n0 + rows_per_group);
  float th[DP], g[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    th[d] = (b < B && d < D) ? theta[(size_t)b * D + d] : 0.f;
    g[d] = 0.f;
  }
  float ll = 0.f;
  const int ntiles = (n1 - n0 + kRsRows - 1) / kRsRows;
  auto stage = [&](int tile, int buf) {
    const int r0 = n0 + tile * kRsRows;
    const int nr = min(kRsRows, n1 - r0);
    // X rows are padded to DP floats (16-byte multiples) by the host
    const int n16 = nr * (DP / 4);
    for (int i = threadIdx.x; i < n16; i += kRsThreads)
      __pipeline_memcpy_async(&xs[buf][4 * i], X + (size_t)r0 * DP + 4 * i, 16);
    for (int i = threadIdx.x; i < nr; i += kRsThreads) __pipeline_memcpy_async(&ys[buf][i], y + r0 + i, 4);
    __pipeline_commit();
  };
  if (ntiles > 0) stage(0, 0);
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      stage(t + 1, buf ^ 1);
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
    const int nr = min(kRsRows, n1 - (n0 + t * kRsRows));
#pragma unroll 1
    for (int r = 0; r < nr; ++r) {
      const float4* xr = reinterpret_cast<const float4*>(&xs[buf][r * DP]);
      float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;
#pragma unroll
      for (int q = 0; q < DP / 4; ++q) {
        const float4 v = xr[q];
        z0 = fmaf(v.x, th[4 * q + 0], z0);
        z1 = fmaf(v.y, th[4 * q + 1], z1);
        z2 = fmaf(v.z, th[4 * q + 2], z2);
        z3 = fmaf(v.w, th[4 * q + 3], z3);
      }
      const float z = (z0 + z1) + (z2 + z3);
      const float yn = ys[buf][r];
      const float e = expf(-fabsf(z));                     // accurate forms: this log-likelihood feeds the MH test
      const float rr = 1.0f / (1.0f + e);
      const float sg = z >= 0.f ? rr : e * rr;
      ll += yn * z - (log1pf(e) + fmaxf(z, 0.f));          // bernoulli.py:119-135
      const float w = yn - sg;
#pragma unroll
      for (int q = 0; q < DP / 4; ++q) {
        const float4 v = xr[q];
        g[4 * q + 0] = fmaf(v.x, w, g[4 * q + 0]);
        g[4 * q + 1] = fmaf(v.y, w, g[4 * q + 1]);
        g[4 * q + 2] = fmaf(v.z, w, g[4 * q + 2]);
        g[4 * q + 3] = fmaf(v.w, w, g[4 * q + 3]);
      }
    }
    __syncthreads();
  }
  if (b < B) {
    float* out = part_g + ((size_t)rg * B + b) * D;
#pragma unroll
    for (int d = 0; d < DP; ++d)
      if (d < D) out[d] = g[d];
    part_ll[(size_t)rg * B + b] = ll;
  }
}

// out[b, 0:D] = sum_rg part_g[rg,b,:], out[b, D] = sum_rg part_ll[rg,b]   (packed [B, D+1] for ONE all-reduce)
__global__ void rowshard_reduce_kernel(const float* part_g, const float* part_ll, int RG, int B, int D,
                                       float* out /*[B, D+1]*/) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tot = (size_t)B * (D + 1);
  if (i >= tot) return;
  const int b = (int)(i / (D + 1)), d = (int)(i - (size_t)b * (D + 1));
  float s = 0.f;
  if (d < D) {
    for (int r = 0; r < RG; ++r) s += part_g[((size_t)r * B + b) * D + d];
  } else {
    for (int r = 0; r < RG; ++r) s += part_ll[(size_t)r * B + b];
  }
  out[i] = s;
}

// packed [B, D+1] (all-reduced likelihood gradient | log-likelihood) + N(0,1) prior on theta
//   grad = -theta + G ; logp = sum_d(-theta^2/2 - log sqrt(2 pi)) + loglik   (logistic_regression.py:88-103,
//   bayesian_model.py:100-102)
__global__ void rowshard_finish_kernel(const float* packed, const float* theta, int B, int D, float* grad,
                                       float* logp) {
  const int b = blockIdx.x;
  float part = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float t = theta[(size_t)b * D + d];
    grad[(size_t)b * D + d] = -t + packed[(size_t)b * (D + 1) + d];
    part += -0.5f * t * t - kHalfLog2Pi;
  }
  part = warp_sum(part);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sh[k];
    logp[b] = s + packed[(size_t)b * (D + 1) + D];
  }
}

// Lock-step leapfrog pieces on [B,D] arrays (leapfrog_integrator.py:280-309,330-355):
//   mode 0: v = m + (0.5 eps) g ; x += eps v            (first half kick + drift)
//   mode 1: v += eps g ; x += eps v                     (full kick with the new gradient + next drift)
//   mode 2: v += eps g ; m_out = v - (0.5 eps) g        (last kick, back to integer-step momentum)
__global__ void lockstep_leapfrog_kernel(int mode, int B, int D, const float* step, int step_kind, float* v,
                                         float* x, const float* g, const float* m_in, float* m_out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * D) return;
  const int b = (int)(i / D), d = (int)(i - (size_t)b * D);
  const float eps = step_kind == 0 ? step[0] : (step_kind == 1 ? step[d] : step[b]);
  if (mode == 0) {
    const float vv = m_in[i] + (0.5f * eps) * g[i];
    v[i] = vv;
    x[i] = x[i] + eps * vv;
  } else if (mode == 1) {
    const float vv = v[i] + eps * g[i];
    v[i] = vv;
    x[i] = x[i] + eps * vv;
  } else {
    const float vv = v[i] + eps * g[i];
    v[i] = vv;
    m_out[i] = vv - (0.5f * eps) * g[i];
  }
}


// ---- Metropolis-Hastings step of a lock-step HMC transition, one kernel: log_acceptance_correction
// (hmc.py:862-875), log_accept_ratio with safe_sum (metropolis_hastings.py:204-215, util.py:205-235), the uniform draw at
// counter = global chain index, accept iff log u < ratio (:221-227), and mcmc_util.choose (util.py:103-164) of the state
// and of every per-chain field of the kernel results.  One block per chain.
struct MhFinishArgs {
  int B, D, chain_offset, layout;
  unsigned long long B_global;
  Key key;
  const float *m0, *m1, *x0, *lp0, *g0, *x1, *lp1, *g1, *pm0, *pm1, *pcorr;
  float *x_out, *lp_out, *g_out, *m0_out, *m1_out, *corr_acc, *corr, *ratio;
  unsigned char* accepted;
};

__global__ void mh_finish_kernel(const MhFinishArgs a) {
  const int b = blockIdx.x;
  const size_t row = (size_t)b * a.D;
  float s0 = 0.f, s1 = 0.f;
  for (int d = threadIdx.x; d < a.D; d += blockDim.x) {
    const float u = a.m0[row + d], v = a.m1[row + d];
    s0 = fmaf(u, u, s0);
    s1 = fmaf(v, v, s1);
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  __shared__ float sh[2][4];
  __shared__ int acc_s;
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s0; sh[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float k0 = (sh[0][0] + sh[0][1]) + (sh[0][2] + sh[0][3]), k1 = (sh[1][0] + sh[1][1]) + (sh[1][2] + sh[1][3]);
    const float corr = 0.5f * finite_or_neginf(k0 + (-k1));
    const float ratio = finite_or_neginf((a.lp1[b] + (-a.lp0[b])) + corr);
    const float u = uniform_from_bits(bits_at(a.key, (uint64_t)a.chain_offset + (uint64_t)b, a.B_global, a.layout), 0.f, 1.f);
    const bool acc = logf(u) < ratio;
    a.corr[b] = corr;
    a.ratio[b] = ratio;
    a.accepted[b] = acc ? 1 : 0;
    a.lp_out[b] = acc ? a.lp1[b] : a.lp0[b];
    a.corr_acc[b] = acc ? corr : a.pcorr[b];
    acc_s = acc ? 1 : 0;
  }
  __syncthreads();
  const bool acc = acc_s != 0;
  for (int d = threadIdx.x; d < a.D; d += blockDim.x) {
    a.x_out[row + d] = acc ? a.x1[row + d] : a.x0[row + d];
    a.g_out[row + d] = acc ? a.g1[row + d] : a.g0[row + d];
    a.m0_out[row + d] = acc ? a.m0[row + d] : a.pm0[row + d];
    a.m1_out[row + d] = acc ? a.m1[row + d] : a.pm1[row + d];
  }
}

}  // namespace pb2

using namespace pb2;

extern "C" {

int pb2_rowshard_logistic_grad(pb2_ctx* ctx, const float* d_X, const float* d_y, int N, int D, int DP,
                               const float* d_theta, int B, float* d_packed) {
  if (!ctx || !d_X || !d_y || !d_theta || !d_packed || N < 0 || B < 1 || D < 1 || DP < D)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_rowshard_logistic_grad: bad argument");
  cudaSetDevice(ctx->device);
  const int chain_tiles = (B + kRsThreads - 1) / kRsThreads;
  int RG = std::max(1, (2 * ctx->num_sms) / chain_tiles);
  RG = std::min(RG, std::max(1, (N + kRsRows - 1) / kRsRows));
  int rows_per_group = (N + RG - 1) / RG;
  rows_per_group = ((rows_per_group + kRsRows - 1) / kRsRows) * kRsRows;
  RG = std::max(1, (N + rows_per_group - 1) / rows_per_group);
  const size_t need = sizeof(float) * ((size_t)RG * B * D + (size_t)RG * B);
  if (need > ctx->sched_bytes) {
    if (ctx->d_sched) cudaFree(ctx->d_sched);
    ctx->d_sched = nullptr;
    ctx->sched_bytes = 0;
    if (int rc = check_cuda(ctx, cudaMalloc((void**)&ctx->d_sched, need), "cudaMalloc(rowshard partials)")) return rc;
    ctx->sched_bytes = need;
  }
  float* part_g = reinterpret_cast<float*>(ctx->d_sched);
  float* part_ll = part_g + (size_t)RG * B * D;
  dim3 grid(chain_tiles, RG);
  switch (DP) {
    case 32: rowshard_logistic_kernel<32><<<grid, kRsThreads, 0, ctx->stream>>>(d_X, d_y, N, D, d_theta, B, rows_per_group, part_g, part_ll); break;
    case 64: rowshard_logistic_kernel<64><<<grid, kRsThreads, 0, ctx->stream>>>(d_X, d_y, N, D, d_theta, B, rows_per_group, part_g, part_ll); break;
    case 100: rowshard_logistic_kernel<100><<<grid, kRsThreads, 0, ctx->stream>>>(d_X, d_y, N, D, d_theta, B, rows_per_group, part_g, part_ll); break;
    default: return set_error(ctx, PB2_ERR_UNSUPPORTED, "pb2_rowshard_logistic_grad: padded row length DP must be 32, 64 or 100");
  }
  ctx->launches += 1;
  if (int rc = check_cuda(ctx, cudaGetLastError(), "rowshard_logistic_kernel")) return rc;
  const size_t tot = (size_t)B * (D + 1);
  rowshard_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(part_g, part_ll, RG, B, D, d_packed);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "rowshard_reduce_kernel");
}

int pb2_rowshard_logistic_finish(pb2_ctx* ctx, const float* d_packed, const float* d_theta, int B, int D,
                                 float* d_grad, float* d_logp) {
  if (!ctx || !d_packed || !d_theta || !d_grad || !d_logp || B < 1 || D < 1)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_rowshard_logistic_finish: bad argument");
  cudaSetDevice(ctx->device);
  rowshard_finish_kernel<<<B, 128, 0, ctx->stream>>>(d_packed, d_theta, B, D, d_grad, d_logp);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "rowshard_finish_kernel");
}

int pb2_lockstep_leapfrog(pb2_ctx* ctx, int mode, int B, int D, const float* d_step, int step_kind, float* d_v,
                          float* d_x, const float* d_g, const float* d_m_in, float* d_m_out) {
  if (!ctx || mode < 0 || mode > 2 || !d_step || !d_v || !d_g || B < 1 || D < 1 || (mode != 2 && !d_x) ||
      (mode == 0 && !d_m_in) || (mode == 2 && !d_m_out) || step_kind < 0 || step_kind > 2)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_lockstep_leapfrog: bad argument");
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)B * D;
  lockstep_leapfrog_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(mode, B, D, d_step, step_kind, d_v,
                                                                               d_x, d_g, d_m_in, d_m_out);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "lockstep_leapfrog_kernel");
}

int pb2_hmc_mh_finish(pb2_ctx* ctx, int B, int D, int B_global, int chain_offset, int rng_layout,
                      const uint32_t accept_key[2], const float* d_m0, const float* d_m1, const float* d_x0,
                      const float* d_lp0, const float* d_g0, const float* d_x1, const float* d_lp1, const float* d_g1,
                      const float* d_prev_m0, const float* d_prev_m1, const float* d_prev_corr, float* d_x_out,
                      float* d_lp_out, float* d_g_out, float* d_m0_out, float* d_m1_out, float* d_corr_acc_out,
                      float* d_corr, float* d_log_accept_ratio, unsigned char* d_is_accepted) {
  if (!ctx || B < 1 || D < 1 || B_global < B || chain_offset < 0 || !accept_key || !d_m0 || !d_m1 || !d_x0 || !d_lp0 ||
      !d_g0 || !d_x1 || !d_lp1 || !d_g1 || !d_prev_m0 || !d_prev_m1 || !d_prev_corr || !d_x_out || !d_lp_out || !d_g_out ||
      !d_m0_out || !d_m1_out || !d_corr_acc_out || !d_corr || !d_log_accept_ratio || !d_is_accepted ||
      rng_layout < PB2_LAYOUT_PARTITIONABLE || rng_layout > PB2_LAYOUT_PHILOX)
    return set_error(ctx, PB2_ERR_INVALID, "pb2_hmc_mh_finish: bad argument");
  cudaSetDevice(ctx->device);
  MhFinishArgs a{B, D, chain_offset, rng_layout, (unsigned long long)B_global, Key{accept_key[0], accept_key[1]},
                 d_m0, d_m1, d_x0, d_lp0, d_g0, d_x1, d_lp1, d_g1, d_prev_m0, d_prev_m1, d_prev_corr,
                 d_x_out, d_lp_out, d_g_out, d_m0_out, d_m1_out, d_corr_acc_out, d_corr, d_log_accept_ratio, d_is_accepted};
  mh_finish_kernel<<<B, 128, 0, ctx->stream>>>(a);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "mh_finish_kernel");
}

}  // extern "C"
