// Threefry-2x32-20 counter RNG with jax.random key/counter conventions (+ Philox-4x32-10 as a second bit generator).
//
// Replaces, on the device, what tfp/internal/samplers.py:200-368 calls on the JAX
// substrate (jax.random.split / fold_in / uniform / normal / randint via
// tfp/internal/backend/numpy/random_generators.py:151-158,278-302).  uint32 streams
// are bit-exact w.r.t. that scheme; float transforms follow the same formulas.
#pragma once
#include "pb2_compat.cuh"

#define PB2_HD __host__ __device__ __forceinline__

namespace pb2 {

enum : int { kLayoutPartitionable = 0, kLayoutOriginal = 1, kLayoutPhilox = 2 };

struct Key {
  uint32_t k0, k1;
};

PB2_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

PB2_HD void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t& o0,
                         uint32_t& o1) {
  const uint32_t ks0 = k0, ks1 = k1, ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += ks0;
  x1 += ks1;
#define PB2_TF_R(r) \
  x0 += x1;         \
  x1 = rotl32(x1, r); \
  x1 ^= x0;
  PB2_TF_R(13) PB2_TF_R(15) PB2_TF_R(26) PB2_TF_R(6)
  x0 += ks1; x1 += ks2 + 1u;
  PB2_TF_R(17) PB2_TF_R(29) PB2_TF_R(16) PB2_TF_R(24)
  x0 += ks2; x1 += ks0 + 2u;
  PB2_TF_R(13) PB2_TF_R(15) PB2_TF_R(26) PB2_TF_R(6)
  x0 += ks0; x1 += ks1 + 3u;
  PB2_TF_R(17) PB2_TF_R(29) PB2_TF_R(16) PB2_TF_R(24)
  x0 += ks1; x1 += ks2 + 4u;
  PB2_TF_R(13) PB2_TF_R(15) PB2_TF_R(26) PB2_TF_R(6)
  x0 += ks2; x1 += ks0 + 5u;
#undef PB2_TF_R
  o0 = x0;
  o1 = x1;
}

// Philox-4x32-10 (Salmon et al., SC'11): the generator of tf.random.stateless_* on the reference's TF substrate
// (samplers.py:249-250,324-325,367-368).  Word `w` of the block at 128-bit counter (c0, c1, 0, 0).
PB2_HD uint32_t philox4x32_10_word(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, int w) {
  uint32_t c2 = 0u, c3 = 0u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)c0 * 0xD2511F53ull, p1 = (uint64_t)c2 * 0xCD9E8D57ull;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return w == 0 ? c0 : (w == 1 ? c1 : (w == 2 ? c2 : c3));
}

// out of line: keeps the 10 Philox rounds out of the callers' instruction stream (register allocation of the tile kernels)
__host__ __device__ __noinline__ inline uint32_t philox_bits_at(Key k, uint64_t idx) {
  const uint64_t blk = idx >> 2;
  return philox4x32_10_word(k.k0, k.k1, (uint32_t)blk, (uint32_t)(blk >> 32), (int)(idx & 3));
}

// Element `idx` of a flat draw of `n` uint32 (row-major over the requested shape).
PB2_HD uint32_t bits_at(Key k, uint64_t idx, uint64_t n, int layout) {
  uint32_t o0, o1;
  if (layout == kLayoutPartitionable) {   // the default first: the hot kernels draw through this function
    threefry2x32(k.k0, k.k1, (uint32_t)(idx >> 32), (uint32_t)idx, o0, o1);
    return o0 ^ o1;
  }
  if (layout == kLayoutPhilox) return philox_bits_at(k, idx);
  const uint64_t half = (n + 1) >> 1;  // counters iota(n) padded to even, split in halves
  if (idx < half) {
    uint64_t c1 = half + idx;
    threefry2x32(k.k0, k.k1, (uint32_t)idx, (c1 >= n) ? 0u : (uint32_t)c1, o0, o1);
    return o0;
  }
  threefry2x32(k.k0, k.k1, (uint32_t)(idx - half), (uint32_t)idx, o0, o1);
  return o1;
}

// Child `j` of split(key, n).
PB2_HD Key split_at(Key k, uint32_t j, uint32_t n, int layout) {
  Key out;
  if (layout == kLayoutPartitionable) {
    threefry2x32(k.k0, k.k1, 0u, j, out.k0, out.k1);
    return out;
  }
  out.k0 = bits_at(k, 2ull * j, 2ull * n, layout);
  out.k1 = bits_at(k, 2ull * j + 1, 2ull * n, layout);
  return out;
}

PB2_HD Key fold_in(Key k, uint32_t data) {
  Key out;
  threefry2x32(k.k0, k.k1, 0u, data, out.k0, out.k1);
  return out;
}

// jax.random.uniform float32 on [0,1): mantissa trick.
__device__ __forceinline__ float u01_from_bits(uint32_t b) {
  return __uint_as_float((b >> 9) | 0x3F800000u) - 1.0f;
}

__device__ __forceinline__ float uniform_from_bits(uint32_t b, float lo, float hi) {
  float f = u01_from_bits(b);
  // floats * (maxval - minval) + minval evaluated without contraction, then max(minval, .)
  return fmaxf(lo, __fadd_rn(__fmul_rn(f, hi - lo), lo));
}

// Single-precision erfinv (Giles) -- the polynomial XLA evaluates for f32.
__device__ __forceinline__ float erfinv_f32(float x) {
  float w = -log1pf(-x * x);
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, w, 3.43273939e-07f);
    p = fmaf(p, w, -3.5233877e-06f);
    p = fmaf(p, w, -4.39150654e-06f);
    p = fmaf(p, w, 0.00021858087f);
    p = fmaf(p, w, -0.00125372503f);
    p = fmaf(p, w, -0.00417768164f);
    p = fmaf(p, w, 0.246640727f);
    p = fmaf(p, w, 1.50140941f);
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, w, 0.000100950558f);
    p = fmaf(p, w, 0.00134934322f);
    p = fmaf(p, w, -0.00367342844f);
    p = fmaf(p, w, 0.00573950773f);
    p = fmaf(p, w, -0.0076224613f);
    p = fmaf(p, w, 0.00943887047f);
    p = fmaf(p, w, 1.00167406f);
    p = fmaf(p, w, 2.83297682f);
  }
  float r = p * x;
  return (fabsf(x) == 1.0f) ? copysignf(INFINITY, x) : r;
}

// jax.random.normal float32: sqrt(2) * erfinv(uniform(nextafter(-1, 0), 1)).
__device__ __forceinline__ float normal_from_bits(uint32_t b) {
  const float lo = -0.99999994f;  // nextafterf(-1, 0)
  float u = uniform_from_bits(b, lo, 1.0f);
  return 1.41421356237309504880f * erfinv_f32(u);
}

}  // namespace pb2
