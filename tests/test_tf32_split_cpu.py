"""CPU check of the 3xTF32 operand split used by the tcgen05 kernels (pb2_tile.cuh stage_a, pb2_logistic_tc.cu):
    hi = (bits(x) + 0x1000) & 0xffffe000 ;  lo = tf32(x - hi)   (integer add-and-mask instead of cvt.rna.tf32.f32)
The tensor core multiplies X P as Xhi Phi + Xlo Phi + Xhi Plo; with both operands split this way the dropped
term Xlo Plo and the rounding of lo are O(2^-21) relative, i.e. FP32-accurate contractions."""
import numpy as np


def tf32_round(x):
  b = np.asarray(x, np.float32).view(np.uint32)
  return ((b + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)


def test_integer_rounding_is_round_to_nearest_tf32():
  rng = np.random.default_rng(0)
  x = (rng.standard_normal(200000) * np.exp(rng.uniform(-30, 30, 200000))).astype(np.float32)
  hi = tf32_round(x)
  assert np.all((hi.view(np.uint32) & np.uint32(0x1fff)) == 0)            # 10 explicit mantissa bits
  ulp = np.float32(2.0) ** (np.floor(np.log2(np.abs(x))) - 10)           # tf32 spacing at x
  assert np.all(np.abs(hi.astype(np.float64) - x) <= 0.5 * ulp.astype(np.float64) * (1 + 1e-7))
  assert np.all(np.sign(hi) == np.sign(x))


def test_hi_plus_lo_reconstructs_to_2_pow_minus_21():
  rng = np.random.default_rng(1)
  x = rng.standard_normal(100000).astype(np.float32) * 37
  hi = tf32_round(x)
  lo = tf32_round(x - hi)                                                # x - hi is exact in float32
  assert np.all((x - hi).astype(np.float64) == x.astype(np.float64) - hi.astype(np.float64))
  rel = np.abs(hi.astype(np.float64) + lo.astype(np.float64) - x) / np.abs(x)
  assert rel.max() < 2.0 ** -21


def test_three_pass_contraction_is_fp32_accurate():
  """Xhi Phi + Xlo Phi + Xhi Plo with exact products / float32 accumulation (the tensor core's arithmetic up to its
  accumulator rounding) vs float64, against a plain float32 matmul: same error level."""
  rng = np.random.default_rng(2)
  X = rng.standard_normal((256, 100)).astype(np.float32)
  P = (rng.standard_normal((100, 100)) * np.exp(rng.uniform(-3, 3, (100, 100)))).astype(np.float32)
  Xh, Ph = tf32_round(X), tf32_round(P)
  Xl, Pl = tf32_round(X - Xh), tf32_round(P - Ph)
  ref = X.astype(np.float64) @ P.astype(np.float64)
  split = (Xh.astype(np.float64) @ Ph + Xl.astype(np.float64) @ Ph + Xh.astype(np.float64) @ Pl).astype(np.float32)
  plain = X @ P
  scale = np.abs(ref).max(1, keepdims=True)
  e_split = np.max(np.abs(split - ref) / scale)
  e_plain = np.max(np.abs(plain - ref) / scale)
  assert e_split < 4 * max(e_plain, 1e-7), (e_split, e_plain)
