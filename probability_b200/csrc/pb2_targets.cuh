// Fused log-prob + gradient device functions for the four named targets.
// Each target evaluates (lp, grad) for ONE chain held by a group of threads
// (see pb2_group.cuh); the closed forms restate
//   Eight Schools   tfp/mcmc/eight_schools_hmc.py:41-57 (+ normal.py:182-188)
//   DenseGaussian   inference_gym/targets/ill_conditioned_gaussian.py:77-81,105-106
//   Logistic        inference_gym/targets/logistic_regression.py:88-103, bernoulli.py:119-135
//   StochVol        inference_gym/targets/vectorized_stochastic_volatility.py:233-309,346-356
// with analytic gradients instead of autodiff (tfp/mcmc/internal/util.py:246-308).
#pragma once
#include "pb2_group.cuh"

namespace pb2 {

constexpr float kHalfLog2Pi = 0.91893853320467274178f;

__device__ __forceinline__ float softplusf(float x) {
  return log1pf(expf(-fabsf(x))) + fmaxf(x, 0.f);
}
__device__ __forceinline__ float sigmoidf(float x) {
  float e = expf(-fabsf(x));
  float r = 1.0f / (1.0f + e);
  return x >= 0.f ? r : e * r;
}

// --------------------------------------------------------------------------
// Eight Schools (non-centred): x = [mu, tau, z_0..z_{J-1}], J <= 30, E == 1.
struct EightSchoolsParams {
  const float* y;      // [J] device
  const float* sigma;  // [J] device
  int J;
};

template <class Grp, int E>
struct EightSchoolsT {
  static_assert(E == 1 && !Grp::kIsBlock, "Eight Schools runs warp-per-chain, one element per lane");
  using Params = EightSchoolsParams;
  static constexpr bool kCkptInSmem = true;
  float ys, sig, logsig;
  bool active;
  int J;
  PB2_HOSTFN static size_t cta_smem_floats(const Params&) { return 0; }
  PB2_HOSTFN static size_t group_smem_floats(const Params&) { return 0; }
  __device__ void init_cta(const Params&, float*) {}
  __device__ void init_group(const Params& p, Grp& grp, float*, float*) {
    J = p.J;
    int i = grp.lane - 2;
    active = (i >= 0 && i < J);
    sig = active ? p.sigma[i] : 1.f;
    ys = active ? p.y[i] / sig : 0.f;
    logsig = active ? logf(sig) : 0.f;
  }
  __device__ float logp_grad(Grp& grp, const float (&x)[E], float (&g)[E]) {
    const float mu = grp.bcast(x[0], 0);
    const float tau = grp.bcast(x[0], 1);
    const float e = expf(tau);
    const float z = active ? x[0] : 0.f;
    const float loc = mu + e * z;
    const float r = active ? (ys - loc / sig) : 0.f;  // normal.py:184-185: x/scale - loc/scale
    const float w = r / sig;
    float s[3];
    s[0] = w;
    s[1] = w * z;
    s[2] = active ? ((-0.5f * z * z - kHalfLog2Pi) + (-0.5f * r * r - (kHalfLog2Pi + logsig))) : 0.f;
    grp.template sumN<3>(s);
    const float m10 = mu / 10.f;
    float lp = (-0.5f * m10 * m10 - (kHalfLog2Pi + 2.30258509299404568402f)) +
               (-0.5f * (tau - 5.f) * (tau - 5.f) - kHalfLog2Pi) + s[2];
    float gg;
    if (grp.lane == 0) gg = -mu / 100.f + s[0];
    else if (grp.lane == 1) gg = -(tau - 5.f) + e * s[1];
    else gg = active ? (-z + e * w) : 0.f;
    g[0] = gg;
    return lp;
  }
};

// --------------------------------------------------------------------------
// Dense Gaussian: lp = -1/2 (x-mu)^T P (x-mu) + c ; g = -P (x-mu).  D <= 32*E.
struct DenseGaussianParams {
  const float* P;    // [D, D] device, symmetric
  const float* loc;  // [D] device
  float lognorm;
  int D;
  // (nullable) per-dimension scale s of a diagonally preconditioned run: the kernels then sample u = x / s, i.e. the
  // Gaussian with precision diag(s) P diag(s) and location loc / s (see ScaledT below)
  const float* scale = nullptr;
};
__device__ __forceinline__ float dense_scale_at(const DenseGaussianParams& p, int d) { return p.scale ? p.scale[d] : 1.f; }

template <class Grp, int E>
struct DenseGaussianT {
  static_assert(!Grp::kIsBlock, "dense Gaussian runs (half-)warp-per-chain");
  using Params = DenseGaussianParams;
  static constexpr bool kCkptInSmem = true;
  static constexpr int DP = Grp::G * E;  // padded row length
  const float* sP;                       // CTA-shared [D][DP]
  float* xbuf;                           // per-group [DP]
  float loc[E];
  float lognorm;
  int D;
  PB2_HOSTFN static size_t cta_smem_floats(const Params& p) { return (size_t)p.D * DP; }
  PB2_HOSTFN static size_t group_smem_floats(const Params&) { return DP; }
  __device__ void init_cta(const Params& p, float* cta) {
    for (int i = threadIdx.x; i < p.D * DP; i += blockDim.x) {
      int r = i / DP, c = i - r * DP;
      cta[i] = (c < p.D) ? (dense_scale_at(p, r) * p.P[r * p.D + c]) * dense_scale_at(p, c) : 0.f;
    }
  }
  __device__ void init_group(const Params& p, Grp& grp, float* cta, float* grp_smem) {
    sP = cta;
    xbuf = grp_smem;
    D = p.D;
    lognorm = p.lognorm;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      int d = grp.lane * E + j;
      loc[j] = d < D ? p.loc[d] / dense_scale_at(p, d) : 0.f;
    }
  }
  __device__ float logp_grad(Grp& grp, const float (&x)[E], float (&g)[E]) {
    float xc[E];
#pragma unroll
    for (int j = 0; j < E; ++j) xc[j] = x[j] - loc[j];
    if constexpr (E % 4 == 0) {
#pragma unroll
      for (int q = 0; q < E / 4; ++q)
        *reinterpret_cast<float4*>(xbuf + grp.lane * E + 4 * q) =
            make_float4(xc[4 * q], xc[4 * q + 1], xc[4 * q + 2], xc[4 * q + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < E; ++j) xbuf[grp.lane * E + j] = xc[j];
    }
    grp.sync();
    // two chains per warp: re-converge the halves so ONE LDS of a P row feeds both chains
    if constexpr (Grp::G == 16) __syncwarp();
    float acc[E];
#pragma unroll
    for (int j = 0; j < E; ++j) acc[j] = 0.f;
    const float* prow = sP + grp.lane * E;
    if constexpr (E % 4 == 0) {
      // P symmetric: the column block of row k equals the row block; x_k broadcast as LDS.128 per 4 k
      const int D4 = D & ~3;
      for (int k = 0; k < D4; k += 4) {
        const float4 xk4 = *reinterpret_cast<const float4*>(xbuf + k);
        const float xs[4] = {xk4.x, xk4.y, xk4.z, xk4.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
          for (int q = 0; q < E / 4; ++q) {
            const float4 pv = *reinterpret_cast<const float4*>(prow + (k + kk) * DP + 4 * q);
            acc[4 * q + 0] = fmaf(pv.x, xs[kk], acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(pv.y, xs[kk], acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(pv.z, xs[kk], acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(pv.w, xs[kk], acc[4 * q + 3]);
          }
        }
      }
      for (int k = D4; k < D; ++k) {
        const float xk = xbuf[k];
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
          const float4 pv = *reinterpret_cast<const float4*>(prow + k * DP + 4 * q);
          acc[4 * q + 0] = fmaf(pv.x, xk, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(pv.y, xk, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(pv.z, xk, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(pv.w, xk, acc[4 * q + 3]);
        }
      }
    } else {
      for (int k = 0; k < D; ++k) {
        const float xk = xbuf[k];
#pragma unroll
        for (int j = 0; j < E; ++j) acc[j] = fmaf(prow[k * DP + j], xk, acc[j]);
      }
    }
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      g[j] = -acc[j];
      part = fmaf(xc[j], g[j], part);
    }
    grp.sync();  // xbuf is rewritten by the next call
    return fmaf(0.5f, grp.sum(part), lognorm);
  }
};

// --------------------------------------------------------------------------
// Diagonal preconditioning as a change of variables: HMC / NUTS on x with the diagonal mass matrix M = diag(1 / s^2)
// (momentum ~ N(0, M), velocity s^2 m, kinetic energy 1/2 sum s^2 m^2, U-turn test <rho, velocity>:
// tfp/experimental/mcmc/preconditioned_hmc.py, preconditioned_nuts.py:694-705,964-1030,
// diagonal_mass_matrix_adaptation.py:73) is, step for step, identity-mass HMC / NUTS on u = x / s with momentum s m:
// the same standard-normal draw z is the momentum there, x moves by eps s z in both, and <rho_u, m_u> = <rho_x, s^2 m_x>.
// So the transition code stays untouched and any target gets preconditioning from this wrapper: lp(u) = lp_x(s u),
// grad_u = s grad_x.  (The dense Gaussian folds s into its precision matrix instead, so the tensor-core kernels apply.)
template <class Grp, int E, class Tgt>
struct ScaledT {
  struct Params {
    typename Tgt::Params base;
    const float* scale;  // [D] device
    int D;
  };
  static constexpr bool kCkptInSmem = Tgt::kCkptInSmem;
  Tgt base;
  float sc[E];
  PB2_HOSTFN static size_t cta_smem_floats(const Params& p) { return Tgt::cta_smem_floats(p.base); }
  PB2_HOSTFN static size_t group_smem_floats(const Params& p) { return Tgt::group_smem_floats(p.base); }
  __device__ void init_cta(const Params& p, float* cta) { base.init_cta(p.base, cta); }
  __device__ void init_group(const Params& p, Grp& grp, float* cta, float* gs) {
    base.init_group(p.base, grp, cta, gs);
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int d = grp.lane * E + j;
      sc[j] = d < p.D ? p.scale[d] : 1.f;
    }
  }
  __device__ float logp_grad(Grp& grp, const float (&u)[E], float (&g)[E]) {
    float x[E];
#pragma unroll
    for (int j = 0; j < E; ++j) x[j] = sc[j] * u[j];
    const float lp = base.logp_grad(grp, x, g);
#pragma unroll
    for (int j = 0; j < E; ++j) g[j] = sc[j] * g[j];
    return lp;
  }
};

// --------------------------------------------------------------------------
// TransformedTransitionKernel's target (tfp/mcmc/transformed_kernel.py:86-140,165,369-419): the chain lives in the
// unconstrained u, the wrapped target sees x = b(u), and the density carries the forward log-det-Jacobian:
//   lp(u) = lp_x(b(u)) + sum_d fldj_d(u_d),   grad_u = grad_x b'(u) + fldj'(u)
// with one elementwise bijector per dimension (bijectors/{identity,exp,softplus,sigmoid}.py):
//   Identity        b = u                      fldj = 0
//   Exp             b = e^u                    fldj = u
//   Softplus        b = log(1 + e^u)           fldj = -softplus(-u)
//   Sigmoid(lo,hi)  b = lo + (hi - lo) s(u)    fldj = log(hi - lo) - softplus(-u) - softplus(u)
// An optional diagonal preconditioning scale is applied first (u = scale * w, see ScaledT).
enum { kBijIdentity = 0, kBijExp = 1, kBijSoftplus = 2, kBijSigmoid = 3 };
struct BijectorSpec {
  const int* kind;     // [D] device
  const float* lo;     // [D] device (Sigmoid)
  const float* hi;
};

template <class Grp, int E, class Tgt>
struct TransformedT {
  struct Params {
    typename Tgt::Params base;
    BijectorSpec bij;
    const float* scale;  // nullable
    int D;
  };
  static constexpr bool kCkptInSmem = Tgt::kCkptInSmem;
  Tgt base;
  float sc[E], lo[E], hi[E];
  int kind[E];
  PB2_HOSTFN static size_t cta_smem_floats(const Params& p) { return Tgt::cta_smem_floats(p.base); }
  PB2_HOSTFN static size_t group_smem_floats(const Params& p) { return Tgt::group_smem_floats(p.base); }
  __device__ void init_cta(const Params& p, float* cta) { base.init_cta(p.base, cta); }
  __device__ void init_group(const Params& p, Grp& grp, float* cta, float* gs) {
    base.init_group(p.base, grp, cta, gs);
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int d = grp.lane * E + j;
      const bool in = d < p.D;
      sc[j] = (in && p.scale) ? p.scale[d] : 1.f;
      kind[j] = in ? p.bij.kind[d] : kBijIdentity;
      lo[j] = in ? p.bij.lo[d] : 0.f;
      hi[j] = in ? p.bij.hi[d] : 1.f;
    }
  }
  __device__ float logp_grad(Grp& grp, const float (&w)[E], float (&g)[E]) {
    float x[E], db[E], dj[E];
    float fldj = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const float u = sc[j] * w[j];
      float b = u, d1 = 1.f, d2 = 0.f, lj = 0.f;
      if (kind[j] == kBijExp) {
        b = expf(u); d1 = b; d2 = 1.f; lj = u;
      } else if (kind[j] == kBijSoftplus) {
        b = softplusf(u); d1 = sigmoidf(u); d2 = sigmoidf(-u); lj = -softplusf(-u);
      } else if (kind[j] == kBijSigmoid) {
        const float s = sigmoidf(u), sm = sigmoidf(-u), w_ = hi[j] - lo[j];
        b = lo[j] + w_ * s; d1 = w_ * s * sm; d2 = sm - s; lj = logf(w_) - softplusf(-u) - softplusf(u);
      }
      x[j] = b; db[j] = d1; dj[j] = d2;
      fldj += lj;
    }
    const float lp = base.logp_grad(grp, x, g);
#pragma unroll
    for (int j = 0; j < E; ++j) g[j] = sc[j] * fmaf(g[j], db[j], dj[j]);
    return lp + grp.sum(fldj);
  }
};

// --------------------------------------------------------------------------
// Logistic regression, data resident in shared memory: theta on lanes (E == 1,
// D <= DT <= 32), rows strided over lanes.  X~ already carries the bias column.
struct LogisticParams {
  const float* X;  // [N, D] device row-major
  const float* y;  // [N] device (0/1 as float)
  int N, D;
};

template <class Grp, int E, int DT>
struct LogisticT {
  static_assert(E == 1 && !Grp::kIsBlock && DT <= 32, "logistic runs warp-per-chain, theta on lanes");
  using Params = LogisticParams;
  static constexpr bool kCkptInSmem = true;
  // row stride in floats: a multiple of 4 whose quarter is odd (25 -> 28), so that lanes walking rows read a row as
  // 128-bit loads without bank conflicts (stride = 4 * odd words permutes the 8 x 16 B groups of a quarter-warp)
  static constexpr int RS = (((DT + 3) / 4) | 1) * 4;
  const float* sX;
  const float* sy;
  int N, D;
  PB2_HOSTFN static size_t cta_smem_floats(const Params& p) { return (size_t)p.N * RS + p.N; }
  PB2_HOSTFN static size_t group_smem_floats(const Params&) { return 0; }
  __device__ void init_cta(const Params& p, float* cta) {
    for (int i = threadIdx.x; i < p.N * RS; i += blockDim.x) {
      int r = i / RS, c = i - r * RS;
      cta[i] = (c < p.D) ? p.X[r * p.D + c] : 0.f;
    }
    float* yy = cta + (size_t)p.N * RS;
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) yy[i] = p.y[i];
  }
  __device__ void init_group(const Params& p, Grp&, float* cta, float*) {
    sX = cta;
    sy = cta + (size_t)p.N * RS;
    N = p.N;
    D = p.D;
  }
  __device__ float logp_grad(Grp& grp, const float (&x)[E], float (&g)[E]) {
    float th[RS];
#pragma unroll
    for (int d = 0; d < RS; ++d) th[d] = d < 32 ? __shfl_sync(0xffffffffu, x[0], d & 31) : 0.f;  // lanes >= D hold 0
    // Both contractions of a row run on packed FP32 FMAs (FFMA2: two lanes of a register pair per instruction): the
    // logit is accumulated as (even dims, odd dims) partial sums, the gradient as pairs of adjacent dims.
    constexpr int kPairs = RS / 2;   // the padding columns of a row are zero
    static_assert(kPairs <= 18, "D <= 32");
    float2 acc2[18];
#pragma unroll
    for (int q = 0; q < 18; ++q) acc2[q] = make_float2(0.f, 0.f);
    float ll = 0.f;
    for (int n = grp.lane; n < N; n += 32) {
      const float4* row = reinterpret_cast<const float4*>(sX + n * RS);
      float xr[RS];
#pragma unroll
      for (int q = 0; q < RS / 4; ++q) {
        const float4 v = row[q];
        xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
      }
      float2 z2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < kPairs; ++q)
        z2 = __ffma2_rn(make_float2(xr[2 * q], xr[2 * q + 1]), make_float2(th[2 * q], th[2 * q + 1]), z2);
      const float z = z2.x + z2.y;
      const float yn = sy[n];
      const float e = __expf(-fabsf(z));
      const float r = __fdividef(1.0f, 1.0f + e);
      const float sg = z >= 0.f ? r : e * r;
      ll += yn * z - (__logf(1.0f + e) + fmaxf(z, 0.f));
      const float w = yn - sg;
      const float2 w2 = make_float2(w, w);
#pragma unroll
      for (int q = 0; q < kPairs; ++q) acc2[q] = __ffma2_rn(make_float2(xr[2 * q], xr[2 * q + 1]), w2, acc2[q]);
    }
    float acc[32];
#pragma unroll
    for (int q = 0; q < 16; ++q) { acc[2 * q] = acc2[q].x; acc[2 * q + 1] = acc2[q].y; }
    // transpose-reduce: 31 shuffles leave the total of index l on lane l
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const bool up = (grp.lane & o) != 0;
#pragma unroll
      for (int i = 0; i < o; ++i) {
        const float send = up ? acc[i] : acc[i + o];
        const float keep = up ? acc[i + o] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    const float th_own = x[0];
    const bool own = grp.lane < D;
    g[0] = own ? (-th_own + acc[0]) : 0.f;
    float part = ll + (own ? (-0.5f * th_own * th_own - kHalfLog2Pi) : 0.f);
    return grp.sum(part);
  }
};

// --------------------------------------------------------------------------
// Stochastic volatility (non-centred, unconstrained space): u = [u_phi, m, u_s, z_0..z_{T-1}].
// One CTA per chain; thread `lane` owns d = lane*E + j (time index tau = d - 3), so the
// AR(1) recurrence h_t = phi h_{t-1} + s z_t and its adjoint lambda_t = a_t + phi lambda_{t+1}
// are block-wide scans of affine maps (thread-serial, warp Kogge-Stone, cross-warp via smem).
struct StochVolParams {
  const float* y;  // [T] device, centred returns
  int T;
};

// kFolded = true: state in UNCONSTRAINED coordinates [u_phi, m, u_s, z] with the default event-space bijector
// (Sigmoid(-1, 1), Identity, Softplus, Identity; vectorized_stochastic_volatility.py:346-356) and its log-det-Jacobian
// folded into the target -- what a TransformedTransitionKernel over the constrained model evaluates;
// kFolded = false: the model itself, state [phi, m, s, z] (persistence in (-1, 1), shock scale > 0).
template <class Grp, int E, bool kFolded = true>
struct StochVolT {
  static_assert(Grp::kIsBlock && E >= 4, "stochastic volatility runs CTA-per-chain");
  using Params = StochVolParams;
  static constexpr bool kCkptInSmem = false;
  static constexpr int NW = Grp::G / 32;
  float ysq[E];
  int T;
  float* prm;  // [8] phi, m, s, rs, lp_params
  float* wtf;  // [2*NW] forward warp totals (A,B)
  float* wtr;  // [2*NW] reverse warp totals
  PB2_HOSTFN static size_t cta_smem_floats(const Params&) { return 0; }
  PB2_HOSTFN static size_t group_smem_floats(const Params&) { return 8 + 4 * NW; }
  __device__ void init_cta(const Params&, float*) {}
  __device__ void init_group(const Params& p, Grp& grp, float*, float* gs) {
    T = p.T;
    prm = gs;
    wtf = gs + 8;
    wtr = gs + 8 + 2 * NW;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      int tau = grp.lane * E + j - 3;
      float yv = (tau >= 0 && tau < T) ? p.y[tau] : 0.f;
      ysq[j] = yv * yv;
    }
  }
  __device__ float logp_grad(Grp& grp, const float (&x)[E], float (&g)[E]) {
    const int lane = grp.lane, wl = lane & 31, w = lane >> 5;
    float sg = 0.f, sgm = 0.f, sg3 = 0.f, sgm3 = 0.f;
    if (lane == 0) {
      const float u1 = x[0], mm = x[1], u3 = x[2];
      sg = sigmoidf(u1); sgm = sigmoidf(-u1);
      sg3 = sigmoidf(u3); sgm3 = sigmoidf(-u3);
      const float phi0 = kFolded ? 2.f * sg - 1.f : u1;
      const float s0 = kFolded ? softplusf(u3) : u3;
      const float rs0 = 1.0f / sqrtf(1.f - phi0 * phi0);
      const float b = (phi0 + 1.f) * 0.5f;
      // Beta(20,1.5).log_prob(b) - log 2 ; lbeta(20,1.5) = lgamma(20)+lgamma(1.5)-lgamma(21.5)
      const float lp_phi = 19.f * logf(b) + 0.5f * log1pf(-b) - (-4.63282391111f) - 0.693147180559945f;
      const float m5 = mm / 5.f;
      const float lp_m = -2.75416779828f - log1pf(m5 * m5);           // -log(pi*5)
      const float s2 = s0 * 0.5f;
      const float lp_s = 0.693147180559945f - 1.83787706640935f - log1pf(s2 * s2);  // log2 - log(2 pi)
      const float fldj = kFolded ? (0.693147180559945f - softplusf(-u1) - softplusf(u1)) + (-softplusf(-u3)) : 0.f;
      prm[0] = phi0; prm[1] = mm; prm[2] = s0; prm[3] = rs0;
      prm[4] = lp_phi + lp_m + lp_s + fldj;
    }
    __syncthreads();
    const float phi = prm[0], m = prm[1], s = prm[2], rs = prm[3], lp_params = prm[4];
    // ---- forward scan: h
    float A = 1.f, Bc = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int tau = lane * E + j - 3;
      if (tau >= 0 && tau < T) {
        const float b = (tau == 0) ? s * x[j] * rs : s * x[j];
        Bc = fmaf(phi, Bc, b);
        A *= phi;
      }
    }
    float Ai = A, Bi = Bc;  // inclusive within warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float Ap = __shfl_up_sync(0xffffffffu, Ai, o);
      const float Bp = __shfl_up_sync(0xffffffffu, Bi, o);
      if (wl >= o) { Bi = fmaf(Ai, Bp, Bi); Ai *= Ap; }
    }
    if (wl == 31) { wtf[2 * w] = Ai; wtf[2 * w + 1] = Bi; }
    float Bex = __shfl_up_sync(0xffffffffu, Bi, 1);
    float Aex = __shfl_up_sync(0xffffffffu, Ai, 1);
    if (wl == 0) { Bex = 0.f; Aex = 1.f; }
    __syncthreads();
    // h entering this warp = the composition of the warp totals of warps 0 .. w - 1 applied to 0: an inclusive
    // Kogge-Stone scan of the NW affine maps over the lanes (4 steps) instead of a serial loop of up to NW - 1 steps
    float hw;
    {
      float At = wl < NW ? wtf[2 * wl] : 1.f, Bt = wl < NW ? wtf[2 * wl + 1] : 0.f;
#pragma unroll
      for (int o = 1; o < NW; o <<= 1) {
        const float Ap = __shfl_up_sync(0xffffffffu, At, o);
        const float Bp = __shfl_up_sync(0xffffffffu, Bt, o);
        if (wl >= o) { Bt = fmaf(At, Bp, Bt); At *= Ap; }
      }
      hw = __shfl_sync(0xffffffffu, Bt, w > 0 ? w - 1 : 0);
      if (w == 0) hw = 0.f;
    }
    const float hin = fmaf(Aex, hw, Bex);
    float h[E], a[E];
    float hcur = hin;
    float s_a = 0.f, s_lik = 0.f, s_z = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int tau = lane * E + j - 3;
      h[j] = hcur;  // h_{tau-1}
      a[j] = 0.f;
      if (tau >= 0 && tau < T) {
        const float b = (tau == 0) ? s * x[j] * rs : s * x[j];
        hcur = fmaf(phi, hcur, b);
        const float hm = hcur + m;
        const float y2e = ysq[j] * expf(-hm);
        a[j] = 0.5f * (y2e - 1.f);
        s_a += a[j];
        s_lik += -0.5f * y2e - kHalfLog2Pi - 0.5f * hm;
        s_z += -0.5f * x[j] * x[j] - kHalfLog2Pi;
      }
    }
    // ---- reverse scan: lambda_tau = a_tau + phi * lambda_{tau+1}
    float Ar = 1.f, Br = 0.f;
#pragma unroll
    for (int j = E - 1; j >= 0; --j) {
      const int tau = lane * E + j - 3;
      if (tau >= 0 && tau < T) { Br = fmaf(phi, Br, a[j]); Ar *= phi; }
    }
    float Ari = Ar, Bri = Br;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float Ap = __shfl_down_sync(0xffffffffu, Ari, o);
      const float Bp = __shfl_down_sync(0xffffffffu, Bri, o);
      if (wl + o < 32) { Bri = fmaf(Ari, Bp, Bri); Ari *= Ap; }
    }
    if (wl == 0) { wtr[2 * w] = Ari; wtr[2 * w + 1] = Bri; }
    float Brex = __shfl_down_sync(0xffffffffu, Bri, 1);
    float Arex = __shfl_down_sync(0xffffffffu, Ari, 1);
    if (wl == 31) { Brex = 0.f; Arex = 1.f; }
    __syncthreads();
    // lambda entering this warp from the right = warps NW - 1 .. w + 1 composed: the same scan, mirrored
    float lw;
    {
      float At = wl < NW ? wtr[2 * wl] : 1.f, Bt = wl < NW ? wtr[2 * wl + 1] : 0.f;
#pragma unroll
      for (int o = 1; o < NW; o <<= 1) {
        const float Ap = __shfl_down_sync(0xffffffffu, At, o);
        const float Bp = __shfl_down_sync(0xffffffffu, Bt, o);
        if (wl + o < NW) { Bt = fmaf(At, Bp, Bt); At *= Ap; }
      }
      lw = __shfl_sync(0xffffffffu, Bt, w + 1 < NW ? w + 1 : 0);
      if (w + 1 >= NW) lw = 0.f;
    }
    float lam = fmaf(Arex, lw, Brex);
    float s_lc = 0.f, s_lh = 0.f;
    float lam0 = 0.f, z0 = 0.f;
#pragma unroll
    for (int j = E - 1; j >= 0; --j) {
      const int tau = lane * E + j - 3;
      float gg = 0.f;
      if (tau >= 0 && tau < T) {
        lam = fmaf(phi, lam, a[j]);
        if (tau == 0) {
          s_lc = fmaf(lam, x[j] * rs, s_lc);
          gg = s * lam * rs - x[j];
          lam0 = lam; z0 = x[j];
        } else {
          s_lc = fmaf(lam, x[j], s_lc);
          s_lh = fmaf(lam, h[j], s_lh);
          gg = fmaf(s, lam, -x[j]);
        }
      }
      g[j] = gg;
    }
    float sums[5] = {s_a, s_lik, s_z, s_lc, s_lh};
    grp.template sumN<5>(sums);
    if (lane == 0) {
      const float b = (phi + 1.f) * 0.5f;
      const float m5 = m / 5.f, s2 = s * 0.5f;
      const float d_m = sums[0] - (2.f * m / 25.f) / (1.f + m5 * m5);
      const float d_s = sums[3] - s2 / (1.f + s2 * s2);
      const float d_phi = sums[4] + lam0 * s * z0 * phi * rs * rs * rs + 0.5f * (19.f / b - 0.5f / (1.f - b));
      g[0] = kFolded ? d_phi * (2.f * sg * sgm) + (sgm - sg) : d_phi;
      g[1] = d_m;
      g[2] = kFolded ? d_s * sg3 + sgm3 : d_s;
    }
    return sums[1] + sums[2] + lp_params;
  }
};

// --------------------------------------------------------------------------
// CENTRED stochastic volatility (inference_gym/targets/stochastic_volatility.py:39-52,68-111): one latent
// log-volatility x_t per time step, x_0 ~ N(m, s / sqrt(1 - phi^2)), x_t ~ N(m + phi (x_{t-1} - m), s),
// y_t ~ N(0, exp(x_t / 2)), the same priors on phi, m, s as the non-centred model above.  No scan: with the residuals
//   e_0 = (x_0 - m) q / s,  e_t = ((x_t - m) - phi (x_{t-1} - m)) / s,  q = sqrt(1 - phi^2),
// the gradient is a 3-point stencil, d/dx_t = -e_t c_t / s + phi e_{t+1} / s + (y_t^2 e^{-x_t} - 1) / 2 (c_0 = q, else 1),
// and the three parameter derivatives are sums over t.  Same CTA-per-chain mapping as StochVolT: state [phi, m, s, x],
// thread `lane` owns elements lane E .. lane E + E - 1, its neighbours' boundary values come by shuffle (and through
// shared memory across warps).  kFolded as above.
template <class Grp, int E, bool kFolded = true>
struct StochVolCenteredT {
  static_assert(Grp::kIsBlock && E >= 4, "stochastic volatility runs CTA-per-chain");
  using Params = StochVolParams;
  static constexpr bool kCkptInSmem = false;
  static constexpr int NW = Grp::G / 32;
  float ysq[E];
  int T;
  float* prm;  // [8] phi, m, s, q, lp_params
  float* wb;   // [2*NW] warp boundary values: last x of warp w | first residual of warp w
  PB2_HOSTFN static size_t cta_smem_floats(const Params&) { return 0; }
  PB2_HOSTFN static size_t group_smem_floats(const Params&) { return 8 + 2 * NW; }
  __device__ void init_cta(const Params&, float*) {}
  __device__ void init_group(const Params& p, Grp& grp, float*, float* gs) {
    T = p.T;
    prm = gs;
    wb = gs + 8;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      int tau = grp.lane * E + j - 3;
      float yv = (tau >= 0 && tau < T) ? p.y[tau] : 0.f;
      ysq[j] = yv * yv;
    }
  }
  __device__ float logp_grad(Grp& grp, const float (&x)[E], float (&g)[E]) {
    const int lane = grp.lane, wl = lane & 31, w = lane >> 5;
    float sg = 0.f, sgm = 0.f, sg3 = 0.f, sgm3 = 0.f;
    if (lane == 0) {
      const float u1 = x[0], mm = x[1], u3 = x[2];
      sg = sigmoidf(u1); sgm = sigmoidf(-u1);
      sg3 = sigmoidf(u3); sgm3 = sigmoidf(-u3);
      const float phi0 = kFolded ? 2.f * sg - 1.f : u1;
      const float s0 = kFolded ? softplusf(u3) : u3;
      const float b = (phi0 + 1.f) * 0.5f;
      const float lp_phi = 19.f * logf(b) + 0.5f * log1pf(-b) - (-4.63282391111f) - 0.693147180559945f;
      const float m5 = mm / 5.f;
      const float lp_m = -2.75416779828f - log1pf(m5 * m5);           // -log(pi*5)
      const float s2 = s0 * 0.5f;
      const float lp_s = 0.693147180559945f - 1.83787706640935f - log1pf(s2 * s2);  // log2 - log(2 pi)
      const float fldj = kFolded ? (0.693147180559945f - softplusf(-u1) - softplusf(u1)) + (-softplusf(-u3)) : 0.f;
      const float q0 = sqrtf(1.f - phi0 * phi0);
      prm[0] = phi0; prm[1] = mm; prm[2] = s0; prm[3] = q0;
      prm[4] = lp_phi + lp_m + lp_s + fldj + logf(q0);               // + log q: the x_0 term's -log(s / q)
    }
    // the x before my first element: the previous thread's last one
    float xprev = __shfl_up_sync(0xffffffffu, x[E - 1], 1);
    if (wl == 31) wb[w] = x[E - 1];
    __syncthreads();
    if (wl == 0) xprev = w > 0 ? wb[w - 1] : 0.f;
    const float phi = prm[0], m = prm[1], s = prm[2], q = prm[3], lp_params = prm[4];
    const float rs_ = 1.f / s, logs = logf(s);
    float e[E];
    float s_lp = 0.f, s_m = 0.f, s_s = 0.f, s_phi = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int tau = lane * E + j - 3;
      float ev = 0.f;
      if (tau >= 0 && tau < T) {
        const float d = x[j] - m;
        const float dp = (j > 0 ? x[j > 0 ? j - 1 : 0] : xprev) - m;
        if (tau == 0) {
          ev = d * q * rs_;
          s_m = fmaf(ev, q, s_m);
          s_phi = fmaf(ev * d, phi / q, s_phi);
        } else {
          ev = (d - phi * dp) * rs_;
          s_m = fmaf(ev, 1.f - phi, s_m);
          s_phi = fmaf(ev, dp, s_phi);
        }
        s_s += ev * ev - 1.f;
        const float y2e = ysq[j] * expf(-x[j]);
        s_lp += (-0.5f * ev * ev - kHalfLog2Pi - logs) + (-0.5f * y2e - kHalfLog2Pi - 0.5f * x[j]);
        g[j] = 0.5f * (y2e - 1.f);   // the likelihood part; the stencil is added below
      } else {
        g[j] = 0.f;
      }
      e[j] = ev;
    }
    // the residual after my last element: the next thread's first one
    float enext = __shfl_down_sync(0xffffffffu, e[0], 1);
    if (wl == 0) wb[NW + w] = e[0];
    __syncthreads();
    if (wl == 31) enext = w + 1 < NW ? wb[NW + w + 1] : 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int tau = lane * E + j - 3;
      if (tau >= 0 && tau < T) {
        const float en = j + 1 < E ? e[j + 1 < E ? j + 1 : 0] : enext;   // 0 past the end of the series
        const float c = tau == 0 ? q : 1.f;
        g[j] += (phi * en - e[j] * c) * rs_;
      }
    }
    float sums[4] = {s_lp, s_m, s_s, s_phi};
    grp.template sumN<4>(sums);
    if (lane == 0) {
      const float b = (phi + 1.f) * 0.5f;
      const float m5 = m / 5.f, s2 = s * 0.5f;
      const float d_m = sums[1] * rs_ - (2.f * m / 25.f) / (1.f + m5 * m5);
      const float d_s = sums[2] * rs_ - s2 / (1.f + s2 * s2);
      const float d_phi = sums[3] * rs_ - phi / (q * q) + 0.5f * (19.f / b - 0.5f / (1.f - b));
      g[0] = kFolded ? d_phi * (2.f * sg * sgm) + (sgm - sg) : d_phi;
      g[1] = d_m;
      g[2] = kFolded ? d_s * sg3 + sgm3 : d_s;
    }
    return sums[0] + lp_params;
  }
};

}  // namespace pb2
