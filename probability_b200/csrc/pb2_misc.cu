// Small kernels around the chain kernels: key-schedule expansion, bulk RNG draws,
// dual-averaging step-size adaptation, ESS and R-hat cross-chain reductions.
#include <cmath>
#include <algorithm>
#include "pb2_internal.h"

namespace pb2 {

// ---------------------------------------------------------------- key schedules
// HMC transition t (metropolis_hastings.py:183, hmc.py:684-685):
//   prop, acc = split(seed_t); part_keys = split(prop, n_parts)
// layout: [n_parts x 2] part keys | [2] acc key
__global__ void hmc_sched_kernel(const uint32_t* step_keys, int T, int n_parts, int layout, uint32_t* out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  Key seed{step_keys[2 * t], step_keys[2 * t + 1]};
  Key prop = split_at(seed, 0, 2, layout);
  Key acc = split_at(seed, 1, 2, layout);
  uint32_t* o = out + (size_t)t * (2 * (n_parts + 1));
  for (int q = 0; q < n_parts; ++q) {
    Key k = split_at(prop, q, n_parts, layout);
    o[2 * q] = k.k0;
    o[2 * q + 1] = k.k1;
  }
  o[2 * n_parts] = acc.k0;
  o[2 * n_parts + 1] = acc.k1;
}

// NUTS transition t (nuts.py:323, :515, :546-558, :808).  One thread per (t, depth j):
//   k_start, k_loop = split(seed_t); ks = split(k_start, n_parts+1)      (momentum keys)
//   key_0 = k_loop; (k_dir, k_sub, k_acc, key_{j+1}) = split(key_j, 4)
//   direction key = split(k_dir, 2)[1]  (jax.random.randint with span 2 uses only the low draw)
//   kk_0 = k_sub; (k_u[i], kk_{i+1}) = split(kk_i, 2), i < 2^j
// layout: [n_parts x 2] | per depth [dir(2) acc(2) sub(2)] | k_u for depth j at word 2*(2^j - 1)
__global__ void nuts_sched_kernel(const uint32_t* step_keys, int T, int n_parts, int max_depth, int layout,
                                  int stride, uint32_t* out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * max_depth) return;
  const int t = idx / max_depth, j = idx - t * max_depth;
  Key seed{step_keys[2 * t], step_keys[2 * t + 1]};
  uint32_t* o = out + (size_t)t * stride;
  if (j == 0) {
    Key k_start = split_at(seed, 0, 2, layout);
    for (int q = 0; q < n_parts; ++q) {
      Key k = split_at(k_start, q, n_parts + 1, layout);
      o[2 * q] = k.k0;
      o[2 * q + 1] = k.k1;
    }
  }
  Key key = split_at(seed, 1, 2, layout);
  for (int d = 0; d < j; ++d) key = split_at(key, 3, 4, layout);
  Key k_dir = split_at(key, 0, 4, layout);
  Key k_sub = split_at(key, 1, 4, layout);
  Key k_acc = split_at(key, 2, 4, layout);
  Key k_dir2 = split_at(k_dir, 1, 2, layout);
  uint32_t* h = o + 2 * n_parts + 6 * j;
  h[0] = k_dir2.k0; h[1] = k_dir2.k1;
  h[2] = k_acc.k0;  h[3] = k_acc.k1;
  h[4] = k_sub.k0;  h[5] = k_sub.k1;
  uint32_t* ku = o + 2 * n_parts + 6 * max_depth + 2 * ((1 << j) - 1);
  Key kk = k_sub;
  for (int i = 0; i < (1 << j); ++i) {
    Key k_u = split_at(kk, 0, 2, layout);
    Key nx = split_at(kk, 1, 2, layout);
    ku[2 * i] = k_u.k0;
    ku[2 * i + 1] = k_u.k1;
    kk = nx;
  }
}

int launch_hmc_sched(pb2_ctx* ctx, const uint32_t* d_step_keys, int T, int n_parts, int layout, uint32_t* d_out) {
  hmc_sched_kernel<<<(T + 127) / 128, 128, 0, ctx->stream>>>(d_step_keys, T, n_parts, layout, d_out);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "hmc_sched_kernel");
}

int launch_nuts_sched(pb2_ctx* ctx, const uint32_t* d_step_keys, int T, int n_parts, int max_depth, int layout,
                      uint32_t* d_out) {
  int n = T * max_depth;
  nuts_sched_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(d_step_keys, T, n_parts, max_depth, layout,
                                                          nuts_sched_stride(n_parts, max_depth), d_out);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "nuts_sched_kernel");
}

// ---------------------------------------------------------------- bulk RNG
enum { kFillBits = 0, kFillUniform = 1, kFillNormal = 2, kFillRandint = 3 };

__global__ void rng_fill_kernel(Key key, Key key_hi, long long n, int layout, int what, float lo, float hi,
                                int ilo, int ihi, void* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint32_t b = bits_at(key, (uint64_t)i, (uint64_t)n, layout);
    if (what == kFillBits) {
      ((uint32_t*)out)[i] = b;
    } else if (what == kFillUniform) {
      ((float*)out)[i] = uniform_from_bits(b, lo, hi);
    } else if (what == kFillNormal) {
      ((float*)out)[i] = normal_from_bits(b);
    } else {
      // jax.random.randint: (hi_bits % span) * ((2^16 % span)^2 % span) + lo_bits % span, mod span
      uint32_t hb = bits_at(key_hi, (uint64_t)i, (uint64_t)n, layout);
      uint32_t span = (uint32_t)(ihi - ilo);
      uint32_t mult = 65536u % span;
      mult = (uint32_t)(((uint64_t)mult * mult) % span);
      uint32_t off = (uint32_t)((((uint64_t)(hb % span) * mult) + (b % span)) % span);
      ((int32_t*)out)[i] = ilo + (int32_t)off;
    }
  }
}

int launch_rng_fill(pb2_ctx* ctx, Key key, Key key_hi, long long n, int layout, int what, float lo, float hi,
                    int ilo, int ihi, void* out) {
  if (n <= 0) return PB2_OK;
  int blocks = (int)std::min<long long>((n + 255) / 256, (long long)ctx->num_sms * 8);
  rng_fill_kernel<<<blocks, 256, 0, ctx->stream>>>(key, key_hi, n, layout, what, lo, hi, ilo, ihi, out);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "rng_fill_kernel");
}

// ---------------------------------------------------------------- dual averaging
// dual_averaging_step_size_adaptation.py:419-475; log_accept_prob getter
// simple_step_size_adaptation.py:42-48; reduce_logmeanexp math/generic.py:221-274 via
// distribute_lib.reduce_logsumexp :147-162.
constexpr float kDaFixedPointScale = 68719476736.0f;   // 2^36

__global__ void da_partial_kernel(const float* lar, int B, float* partial) {
  // The partial is the SUM of the chains' accept probabilities exp(min(0, log_accept_ratio)) in 64-bit fixed point
  // (2^-36 units; exact for p >= 2^-12, truncated below): integer addition is associative, so the statistic -- and with
  // it the adapted step size -- is bit-identical however the chains are split over launches, blocks or ranks.
  __shared__ unsigned long long sh[32];
  unsigned long long s = 0ull;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    float v = lar[i];
    v = fminf(isfinite(v) ? v : -INFINITY, 0.f);
    s += (unsigned long long)(expf(v) * kDaFixedPointScale);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0ull;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += sh[k];
    uint32_t* out = reinterpret_cast<uint32_t*>(partial);
    out[0] = (uint32_t)tot;
    out[1] = (uint32_t)(tot >> 32);
  }
}

__global__ void da_apply_kernel(const float* partials, int n, double B_global, float* st, float* step_out,
                                float* step_seq_next) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(partials);
  unsigned long long tot = 0ull;
  for (int k = 0; k < n; ++k) tot += (unsigned long long)w[2 * k] | ((unsigned long long)w[2 * k + 1] << 32);
  // exp(log-mean-exp of the log accept probabilities) = their mean
  const float mean_accept = (float)((double)tot / ((double)kDaFixedPointScale * B_global));
  float error_sum = st[0], log_avg = st[1];
  const float log_shrink = st[2];
  const int step_i = (int)st[3];
  const int n_adapt = (int)st[4];
  const float target = st[5], gamma = st[6], t0 = st[7], kappa = st[8];
  float step_size = st[9];
  error_sum = error_sum + target - mean_accept;
  const float t = (float)step_i + 1.f;
  const float log_step = log_shrink - (error_sum * sqrtf(t)) / ((t0 + t) * gamma);
  const float eta = powf(t, -kappa);
  const float new_log_avg = eta * log_step + (1.f - eta) * log_avg;
  const int step_n = step_i + 1;
  if (step_n < n_adapt) step_size = expf(log_step);
  else if (step_n == n_adapt) step_size = expf(new_log_avg);
  if (step_n <= n_adapt) log_avg = new_log_avg;
  st[0] = error_sum;
  st[1] = log_avg;
  st[3] = (float)step_n;
  st[9] = step_size;
  if (step_out) *step_out = step_size;
  if (step_seq_next) *step_seq_next = step_size;
}

int launch_da_partial(pb2_ctx* ctx, const float* d_lar, int B, float* d_partial) {
  da_partial_kernel<<<1, 1024, 0, ctx->stream>>>(d_lar, B, d_partial);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "da_partial_kernel");
}

int launch_da_apply(pb2_ctx* ctx, const float* d_partials, int n, long long B_global, float* d_state,
                    float* d_step_out, float* d_step_seq_next) {
  da_apply_kernel<<<1, 32, 0, ctx->stream>>>(d_partials, n, (double)B_global, d_state, d_step_out, d_step_seq_next);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "da_apply_kernel");
}

__global__ void fill_step_seq_kernel(float* seq, const float* step, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) seq[i] = *step;
}

int launch_fill_step_seq(pb2_ctx* ctx, float* d_seq, const float* d_step, int n) {
  fill_step_seq_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_seq, d_step, n);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "fill_step_seq_kernel");
}

// ---------------------------------------------------------------- diagonal preconditioning / streaming moments
// a[r, d] *= s[d]  (div: /= s[d]) over a [rows, D] array: x <-> u = x / s, grad_x <-> grad_u = s grad_x
__global__ void scale_rows_kernel(float* a, size_t n, int D, const float* s, int div) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float sc = s[i % D];
  a[i] = div ? a[i] / sc : a[i] * sc;
}

int launch_scale_rows(pb2_ctx* ctx, float* d_a, size_t rows, int D, const float* d_s, int div) {
  const size_t n = rows * (size_t)D;
  if (!d_a || n == 0) return PB2_OK;
  scale_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_a, n, D, d_s, div);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "scale_rows_kernel");
}

// Running mean / sum of squared deviations per dimension over the rows of x [rows, D] (Welford within a row segment,
// Chan's pairwise merge across segments and into the running state, both in a fixed order: deterministic).
// experimental/stats/sample_stats.py RunningVariance.update: every row is one new observation.
__global__ void moments_partial_kernel(const float* __restrict__ x, long long rows, int D, long long rows_per_seg,
                                       float* __restrict__ part /*[nseg][3][D]*/) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const long long r0 = (long long)blockIdx.y * rows_per_seg;
  const long long r1 = min(rows, r0 + rows_per_seg);
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float v = x[(size_t)r * D + d];
    n += 1.f;
    const float delta = v - mean;
    mean += delta / n;
    m2 = fmaf(delta, v - mean, m2);
  }
  float* o = part + (size_t)blockIdx.y * 3 * D;
  o[d] = n; o[D + d] = mean; o[2 * D + d] = m2;
}

__global__ void moments_merge_kernel(const float* __restrict__ part, int nseg, int D, float* __restrict__ state) {
  const float n0 = state[0];   // the count is shared by all dimensions: read by everybody before thread 0 rewrites it
  float n_out = n0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float n = n0, mean = state[1 + d], m2 = state[1 + D + d];
    for (int s = 0; s < nseg; ++s) {
      const float* o = part + (size_t)s * 3 * D;
      const float nb = o[d], mb = o[D + d], m2b = o[2 * D + d];
      if (nb > 0.f) {
        const float nt = n + nb, delta = mb - mean;
        mean += delta * (nb / nt);
        m2 += m2b + delta * delta * (n * nb / nt);
        n = nt;
      }
    }
    state[1 + d] = mean;
    state[1 + D + d] = m2;
    n_out = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) state[0] = n_out;
}

int launch_running_moments(pb2_ctx* ctx, const float* d_x, long long rows, int D, float* d_state) {
  if (rows <= 0) return PB2_OK;
  int nseg = (int)std::min<long long>(rows, 4 * (long long)ctx->num_sms);
  const long long rps = (rows + nseg - 1) / nseg;
  nseg = (int)((rows + rps - 1) / rps);
  const size_t need = sizeof(float) * 3 * (size_t)D * nseg;
  if (need > ctx->sched_bytes) {
    if (ctx->d_sched) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_sched); }
    ctx->d_sched = nullptr;
    ctx->sched_bytes = 0;
    if (int rc = check_cuda(ctx, cudaMalloc((void**)&ctx->d_sched, need), "cudaMalloc(moments partials)")) return rc;
    ctx->sched_bytes = need;
  }
  float* part = reinterpret_cast<float*>(ctx->d_sched);
  const int bt = 128;
  moments_partial_kernel<<<dim3((D + bt - 1) / bt, nseg), bt, 0, ctx->stream>>>(d_x, rows, D, rps, part);
  // one block merges: the count lives in state[0] and is read by every dimension before it is rewritten
  moments_merge_kernel<<<1, 1024, 0, ctx->stream>>>(part, nseg, D, d_state);
  ctx->launches += 2;
  return check_cuda(ctx, cudaGetLastError(), "running moments kernels");
}

// ---------------------------------------------------------------- diagnostics
// One thread per series (b,d) of states[N,B,D]; consecutive threads read consecutive
// addresses for a fixed n.  mean / unbiased auto-covariance at lag k (sample_stats.py
// :128-213 computes the same quantity through a zero-padded FFT).
__global__ void series_mean_kernel(const float* x, int N, long long S, float* mean) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc += x[(size_t)n * S + s];
  mean[s] = acc / (float)N;
}

__device__ __forceinline__ float autocov_at(const float* x, int N, long long S, long long s, float mu, int k) {
  float acc = 0.f;
  for (int n = 0; n + k < N; ++n)
    acc = fmaf(x[(size_t)n * S + s] - mu, x[(size_t)(n + k) * S + s] - mu, acc);
  return acc / (float)(N - k);
}

// Per-chain ESS (diagnostic.py:203-336 with cross_chain_dims=None): early exit at the
// first lag failing the threshold / positive-pair rule (equivalent to the cumsum mask).
__global__ void ess_single_kernel(const float* x, int N, long long S, const float* mean, float thr, int use_thr,
                                  int max_lag, int pairs, float* out) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const float mu = mean[s];
  const float c0 = autocov_at(x, N, S, s, mu, 0);
  const float fn = (float)N;
  float wsum = 0.f;
  if (pairs) {
    const int npairs = (max_lag + 1) / 2;
    for (int q = 0; q < npairs; ++q) {
      const int k = 2 * q;
      const float r0 = (k == 0) ? c0 / c0 : autocov_at(x, N, S, s, mu, k) / c0;
      const float r1 = autocov_at(x, N, S, s, mu, k + 1) / c0;
      if (r0 + r1 < 0.f) break;
      wsum += ((fn - (float)k) / fn) * r0 + ((fn - (float)(k + 1)) / fn) * r1;
    }
  } else {
    for (int k = 0; k <= max_lag; ++k) {
      const float r = (k == 0) ? c0 / c0 : autocov_at(x, N, S, s, mu, k) / c0;
      if (use_thr && r < thr) break;
      wsum += ((fn - (float)k) / fn) * r;
    }
  }
  out[s] = fn / (-1.f + 2.f * wsum);
}

// Cross-chain ESS (Vehtari et al. 2021 eq. 10; diagnostic.py:235-269): one CTA per
// dimension d, threads over chains, block-reduced mean auto-covariance per lag.
__global__ void ess_cross_kernel(const float* x, int N, int B, int D, const float* mean, float thr, int use_thr,
                                 int max_lag, int pairs, float* out) {
  const int d = blockIdx.x;
  __shared__ float sh[33];
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int k = 0; k < (int)((blockDim.x + 31) >> 5); ++k) t += sh[k];
    return t;
  };
  const long long S = (long long)B * D;
  // between-chain variance of chain means (unbiased, / (C-1)) and W*(N-1)/N = mean_c autocov_0
  float msum = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) msum += mean[(size_t)b * D + d];
  const float gmean = block_sum(msum) / (float)B;
  float bsum = 0.f, c0sum = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float mu = mean[(size_t)b * D + d];
    bsum += (mu - gmean) * (mu - gmean);
    c0sum += autocov_at(x, N, S, (long long)b * D + d, mu, 0);
  }
  const float between = block_sum(bsum) / (float)(B - 1);
  const float within = block_sum(c0sum) / (float)B;
  const float approx_var = within + between;
  const float fn = (float)N;
  auto rho_at = [&](int k) -> float {
    float a = 0.f;
    if (k == 0) return 1.f - (within - within) / approx_var;
    for (int b = threadIdx.x; b < B; b += blockDim.x)
      a += autocov_at(x, N, S, (long long)b * D + d, mean[(size_t)b * D + d], k);
    const float mac = block_sum(a) / (float)B;
    return 1.f - (within - mac) / approx_var;
  };
  float wsum = 0.f;
  if (pairs) {
    const int npairs = (max_lag + 1) / 2;
    for (int q = 0; q < npairs; ++q) {
      const int k = 2 * q;
      const float r0 = rho_at(k), r1 = rho_at(k + 1);
      if (r0 + r1 < 0.f) break;
      wsum += ((fn - (float)k) / fn) * r0 + ((fn - (float)(k + 1)) / fn) * r1;
    }
  } else {
    for (int k = 0; k <= max_lag; ++k) {
      const float r = rho_at(k);
      if (use_thr && r < thr) break;
      wsum += ((fn - (float)k) / fn) * r;
    }
  }
  if (threadIdx.x == 0) out[d] = (float)B * fn / (-1.f + 2.f * wsum);
}

// R-hat (diagnostic.py:476-567): per (chain-half, d) mean and unbiased variance over the
// sample axis, then reduce over chains per d.  One CTA per d.
__global__ void rhat_kernel(const float* x, int N, int B, int D, int split, float* out) {
  const int d = blockIdx.x;
  __shared__ float sh[33];
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int k = 0; k < (int)((blockDim.x + 31) >> 5); ++k) t += sh[k];
    return t;
  };
  const int halves = split ? 2 : 1;
  const int n = split ? N / 2 : N;  // odd N: last sample dropped
  const int M = B * halves;
  const size_t S = (size_t)B * D;
  float msum = 0.f, vsum = 0.f;
  // chain index m = h*B + b; samples n0 = h*n .. h*n + n
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const int h = m / B, b = m - h * B;
    const float* base = x + (size_t)(h * n) * S + (size_t)b * D + d;
    float mu = 0.f;
    for (int i = 0; i < n; ++i) mu += base[(size_t)i * S];
    mu /= (float)n;
    float v = 0.f;
    for (int i = 0; i < n; ++i) {
      const float dlt = base[(size_t)i * S] - mu;
      v = fmaf(dlt, dlt, v);
    }
    msum += mu;
    vsum += v / (float)(n - 1);
  }
  const float gmean = block_sum(msum) / (float)M;
  const float w = block_sum(vsum) / (float)M;
  float bs = 0.f;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const int h = m / B, b = m - h * B;
    const float* base = x + (size_t)(h * n) * S + (size_t)b * D + d;
    float mu = 0.f;
    for (int i = 0; i < n; ++i) mu += base[(size_t)i * S];
    mu /= (float)n;
    bs += (mu - gmean) * (mu - gmean);
  }
  const float b_div_n = block_sum(bs) / (float)(M - 1);
  if (threadIdx.x == 0) {
    const float fn = (float)n, fm = (float)M;
    const float sigma2 = ((fn - 1.f) / fn) * w + b_div_n;
    out[d] = ((fm + 1.f) / fm) * sigma2 / w - (fn - 1.f) / (fm * fn);
  }
}

int launch_ess(pb2_ctx* ctx, const float* d_states, int N, int B, int D, float thr, int use_thr, int max_lag,
               int pairs, int cross, float* d_mean_scratch, float* d_out) {
  const long long S = (long long)B * D;
  series_mean_kernel<<<(unsigned)((S + 255) / 256), 256, 0, ctx->stream>>>(d_states, N, S, d_mean_scratch);
  ctx->launches += 1;
  if (cross) {
    ess_cross_kernel<<<D, 256, 0, ctx->stream>>>(d_states, N, B, D, d_mean_scratch, thr, use_thr, max_lag, pairs,
                                                d_out);
  } else {
    ess_single_kernel<<<(unsigned)((S + 127) / 128), 128, 0, ctx->stream>>>(d_states, N, S, d_mean_scratch, thr,
                                                                           use_thr, max_lag, pairs, d_out);
  }
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "ess kernels");
}

int launch_rhat(pb2_ctx* ctx, const float* d_states, int N, int B, int D, int split, float* d_out) {
  rhat_kernel<<<D, 256, 0, ctx->stream>>>(d_states, N, B, D, split, d_out);
  ctx->launches += 1;
  return check_cuda(ctx, cudaGetLastError(), "rhat_kernel");
}

}  // namespace pb2
