// "Group" = the set of threads that cooperatively owns ONE chain.
//   WarpG      : 32 lanes  (Eight Schools, dense Gaussian, logistic)
//   BlockG<NW> : a whole CTA of NW warps (stochastic volatility, D = 2519)
// A chain vector of D floats is held in registers, E per thread, blocked ownership:
// thread `lane` owns d = lane*E + j, j < E (zero padded past D).  All reductions
// return the SAME bits to every thread of the group so per-chain control flow
// (accept / U-turn / divergence) stays uniform inside the group.
#pragma once
#include "pb2_compat.cuh"

namespace pb2 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct WarpG {
  static constexpr int G = 32;
  static constexpr bool kIsBlock = false;
  // "cold" NUTS vectors (other end, candidates): registers for the small groups
  static constexpr bool kColdInSmem = false;
  template <int E>
  struct Cold {
    float v[E];
    __device__ __forceinline__ Cold(float*, int, int) {}
    __device__ __forceinline__ float& operator[](int j) { return v[j]; }
    __device__ __forceinline__ const float& operator[](int j) const { return v[j]; }
  };
  int lane;
  float* scratch;  // per-group shared scratch (>= 64 floats), see chain kernels
  __device__ WarpG(float* s) : lane(threadIdx.x & 31), scratch(s) {}
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ float sum(float v) { return warp_sum(v); }
  template <int N>
  __device__ __forceinline__ void sumN(float (&v)[N]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
  }
  __device__ __forceinline__ int bcast_int(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  __device__ __forceinline__ float bcast(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
};

// Two chains per warp, 16 lanes each.  The halves run their own (divergent) control flow and only
// re-converge -- __syncwarp(full mask) inside the target's contraction loop -- so that one
// LDS of the shared operand (the precision matrix row) feeds both chains.
struct HalfWarpG {
  static constexpr int G = 16;
  static constexpr bool kIsBlock = false;
  // "cold" NUTS vectors (other end, candidates): registers for the small groups
  static constexpr bool kColdInSmem = false;
  template <int E>
  struct Cold {
    float v[E];
    __device__ __forceinline__ Cold(float*, int, int) {}
    __device__ __forceinline__ float& operator[](int j) { return v[j]; }
    __device__ __forceinline__ const float& operator[](int j) const { return v[j]; }
  };
  int lane;        // 0..15 inside the half
  unsigned mask;   // lanes of this half
  float* scratch;
  __device__ HalfWarpG(float* s)
      : lane(threadIdx.x & 15), mask((threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu), scratch(s) {}
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  __device__ __forceinline__ float sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, 16);
    return v;
  }
  template <int N>
  __device__ __forceinline__ void sumN(float (&v)[N]) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(mask, v[i], o, 16);
    }
  }
  __device__ __forceinline__ int bcast_int(int v, int src) { return __shfl_sync(mask, v, src, 16); }
  __device__ __forceinline__ float bcast(float v, int src) { return __shfl_sync(mask, v, src, 16); }
};

template <int NW>
struct BlockG {
  static constexpr int G = 32 * NW;
  static constexpr bool kIsBlock = true;
  static constexpr int kMaxN = 8;
  // "cold" NUTS vectors (other end, trajectory / subtree candidates: touched at doubling boundaries and on a
  // multinomial take, not in every leapfrog) live in shared memory, one private slot per thread and element
  // ([vector][element][thread]: conflict-free), so that the hot state fits the register budget of two CTAs per SM
  static constexpr bool kColdInSmem = true;
  template <int E>
  struct Cold {
    float* p;
    __device__ __forceinline__ Cold(float* base, int k, int lane) : p(base + (size_t)k * E * G + lane) {}
    __device__ __forceinline__ float& operator[](int j) { return p[j * G]; }
    __device__ __forceinline__ const float& operator[](int j) const { return p[j * G]; }
  };
  int lane;
  float* scratch;
  float* red;  // [2][kMaxN][NW] double-buffered partials: one __syncthreads per reduction
  int parity;
  __device__ BlockG(float* s, float* r) : lane(threadIdx.x), scratch(s), red(r), parity(0) {}
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  template <int N>
  __device__ __forceinline__ void sumN(float (&v)[N]) {
    static_assert(N <= kMaxN, "too many simultaneous reductions");
    static_assert(NW <= 32 && (NW & (NW - 1)) == 0, "second stage is a butterfly over the warp partials");
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    float* buf = red + parity * (kMaxN * NW);
    parity ^= 1;
    const int w = lane >> 5, wl = lane & 31;
    if (wl == 0) {
#pragma unroll
      for (int i = 0; i < N; ++i) buf[i * NW + w] = v[i];
    }
    __syncthreads();
    // second stage: every warp loads the NW partials (one per lane) and folds them with the same butterfly, so all
    // threads of the CTA end with identical bits (1 LDS + log2(NW) shuffles per value instead of NW loads and adds)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float s = buf[i * NW + (wl & (NW - 1))];
#pragma unroll
      for (int o = NW / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      v[i] = s;
    }
  }
  __device__ __forceinline__ float sum(float v) {
    float a[1] = {v};
    sumN<1>(a);
    return a[0];
  }
  // broadcast through the reduction buffers (value taken from thread `src`)
  __device__ __forceinline__ float bcast(float v, int src) {
    float* buf = red + parity * (kMaxN * NW);
    parity ^= 1;
    if (lane == src) buf[0] = v;
    __syncthreads();
    return buf[0];
  }
  __device__ __forceinline__ int bcast_int(int v, int src) {
    return __float_as_int(bcast(__int_as_float(v), src));
  }
};

}  // namespace pb2
