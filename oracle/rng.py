"""Oracle RNG: threefry2x32 counter RNG with jax.random semantics (TEST INFRASTRUCTURE).

Restates what tfp/internal/samplers.py calls on the JAX substrate:
  split_seed     samplers.py:217-256  -> jax.random.split
  fold_in        samplers.py:200-214  -> jax.random.fold_in
  sanitize_seed  samplers.py:79-172   (salt = sha512(str) mod (2**31-1), folded in)
  normal         samplers.py:308-325  -> random_generators.py:151-158 jax.random.normal
  uniform        samplers.py:356-368  -> random_generators.py:278-302 jax.random.uniform/randint

jax is a third-party dependency (un-pinned, setup.py:110-111) that is absent from
/root/reference; the algorithm below is the published Threefry-2x32-20 (Salmon et
al., SC'11) with jax's key/counter conventions.  Both counter layouts are kept:
  PARTITIONABLE (jax_threefry_partitionable=True, the default of current jax):
      element j of a draw uses counter (hi=0, lo=j) and returns out0 ^ out1;
      split child j is (out0, out1) of counter (0, j).
  ORIGINAL: counters iota(n) padded to even, first half -> x0, second half -> x1,
      output concat(out0, out1)[:n]; split = bits(2n).reshape(n, 2).
A third generator, PHILOX, draws the same flat streams from Philox4x32-10 (key = the uint32[2] key, element j = word
j % 4 of counter block j // 4; split = bits(2n).reshape(n, 2); fold_in and the float transforms are unchanged).
"""
import hashlib

import numpy as np

PARTITIONABLE = 0
ORIGINAL = 1
PHILOX = 2   # Philox4x32-10 bit generator (north_star: "Philox counter RNG keyed from samplers.split_seed")

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_U32 = np.uint32


def _rotl(x, r):
  return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
  """Threefry-2x32, 20 rounds. All args uint32 scalars/arrays (broadcast)."""
  with np.errstate(over='ignore'):
    k0 = np.asarray(k0, _U32)
    k1 = np.asarray(k1, _U32)
    x0 = np.asarray(x0, _U32).copy()
    x1 = np.asarray(x1, _U32).copy()
    ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
    x0 = x0 + ks[0]
    x1 = x1 + ks[1]
    for g in range(5):
      for r in _ROT[g % 2]:
        x0 = x0 + x1
        x1 = _rotl(x1, r)
        x1 = x1 ^ x0
      x0 = x0 + ks[(g + 1) % 3]
      x1 = x1 + ks[(g + 2) % 3] + _U32(g + 1)
    return x0, x1


def philox4x32_10(k0, k1, c0, c1, c2, c3):
  """Philox-4x32, 10 rounds (Salmon et al., SC'11; the generator behind tf.random.stateless_* on the TF substrate,
  samplers.py:249-250,324-325,367-368).  Key = 2 words, counter = 4 words; returns the 4 output words."""
  M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
  W0, W1 = _U32(0x9E3779B9), _U32(0xBB67AE85)
  with np.errstate(over='ignore'):
    k0 = np.asarray(k0, _U32).copy(); k1 = np.asarray(k1, _U32).copy()
    c = [np.asarray(v, _U32).copy() for v in (c0, c1, c2, c3)]
    for _ in range(10):
      p0 = c[0].astype(np.uint64) * M0
      p1 = c[2].astype(np.uint64) * M1
      hi0, lo0 = (p0 >> np.uint64(32)).astype(_U32), (p0 & np.uint64(0xFFFFFFFF)).astype(_U32)
      hi1, lo1 = (p1 >> np.uint64(32)).astype(_U32), (p1 & np.uint64(0xFFFFFFFF)).astype(_U32)
      c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
      k0 = k0 + W0
      k1 = k1 + W1
    return c


def key(seed_int):
  """jax.random.PRNGKey(int) -> uint32[2] = [hi, lo]."""
  seed_int = int(seed_int)
  return np.array([(seed_int >> 32) & 0xFFFFFFFF, seed_int & 0xFFFFFFFF], _U32)


def bits(k, n, layout=PARTITIONABLE):
  """uint32[n] random bits for a flat (row-major) draw of n elements."""
  k = np.asarray(k, _U32)
  n = int(n)
  if n == 0:
    return np.zeros([0], _U32)
  if layout == PARTITIONABLE:
    j = np.arange(n, dtype=np.uint64)
    hi = (j >> np.uint64(32)).astype(_U32)
    lo = (j & np.uint64(0xFFFFFFFF)).astype(_U32)
    o0, o1 = threefry2x32(k[0], k[1], hi, lo)
    return o0 ^ o1
  if layout == PHILOX:
    # element j = word j % 4 of the block at counter (j // 4, 0): 128-bit counter, low words first
    j = np.arange(n, dtype=np.uint64)
    blk = j >> np.uint64(2)
    out = philox4x32_10(k[0], k[1], (blk & np.uint64(0xFFFFFFFF)).astype(_U32), (blk >> np.uint64(32)).astype(_U32),
                        np.zeros(n, _U32), np.zeros(n, _U32))
    return np.choose((j & np.uint64(3)).astype(np.int64), out).astype(_U32)
  cnt = np.arange(n, dtype=_U32)
  if n % 2:
    cnt = np.concatenate([cnt, np.zeros([1], _U32)])
  half = cnt.size // 2
  o0, o1 = threefry2x32(k[0], k[1], cnt[:half], cnt[half:])
  return np.concatenate([o0, o1])[:n]


def split(k, n=2, layout=PARTITIONABLE):
  """jax.random.split -> uint32[n, 2]."""
  k = np.asarray(k, _U32)
  if layout == PARTITIONABLE:
    j = np.arange(n, dtype=_U32)
    o0, o1 = threefry2x32(k[0], k[1], np.zeros_like(j), j)
    return np.stack([o0, o1], axis=-1)
  return bits(k, 2 * n, layout).reshape(n, 2)


def fold_in(k, data):
  """jax.random.fold_in(key, uint32 data) = threefry(key; 0, data)."""
  k = np.asarray(k, _U32)
  o0, o1 = threefry2x32(k[0], k[1], _U32(0), _U32(int(data) & 0xFFFFFFFF))
  return np.array([o0, o1], _U32)


def salt_int(salt):
  """samplers.py:159-162."""
  return int(hashlib.sha512(str(salt).encode('utf-8')).hexdigest(), 16) % (2**31 - 1)


def sanitize_seed(seed, salt=None):
  """samplers.py:79-172 (stateless flavour only): int -> key; optional salt fold_in."""
  if isinstance(seed, (int, np.integer)):
    seed = key(seed)
  seed = np.asarray(seed, _U32)
  if salt is not None:
    seed = fold_in(seed, salt_int(salt))
  return seed


def uniform_from_bits(b, lo=0.0, hi=1.0):
  """jax.random.uniform float32: mantissa trick, then max(lo, f*(hi-lo)+lo)."""
  b = np.asarray(b, _U32)
  f = ((b >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
  lo = np.float32(lo)
  hi = np.float32(hi)
  return np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)


def uniform(k, shape, lo=0.0, hi=1.0, layout=PARTITIONABLE):
  shape = tuple(np.atleast_1d(shape).astype(int)) if np.ndim(shape) else (int(shape),)
  n = int(np.prod(shape)) if len(shape) else 1
  return uniform_from_bits(bits(k, n, layout), lo, hi).reshape(shape)


def erfinv_f32(x):
  """Single-precision erf^-1 (M. Giles, 'Approximating the erfinv function'),
  the polynomial XLA evaluates for float32; w = -log1p(-x*x)."""
  x = np.asarray(x, np.float32)
  f = np.float32
  with np.errstate(divide='ignore', invalid='ignore'):
    w = -np.log1p(-x * x).astype(np.float32)
    lt = w < f(5.0)
    wa = w - f(2.5)
    wb = np.sqrt(np.maximum(w, f(0))).astype(np.float32) - f(3.0)
    ca = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
          0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)
    cb = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
          0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)
    pa = np.full_like(x, f(ca[0]))
    for c in ca[1:]:
      pa = (f(c) + pa * wa).astype(np.float32)
    pb = np.full_like(x, f(cb[0]))
    for c in cb[1:]:
      pb = (f(c) + pb * wb).astype(np.float32)
    p = np.where(lt, pa, pb)
    out = (p * x).astype(np.float32)
    out = np.where(np.abs(x) == f(1.0), np.copysign(f(np.inf), x), out)
  return out.astype(np.float32)


def normal_from_bits(b):
  """jax.random.normal float32: sqrt(2)*erfinv(uniform(nextafter(-1,0), 1))."""
  lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
  u = uniform_from_bits(b, lo, 1.0)
  return (np.float32(np.sqrt(2.0)) * erfinv_f32(u)).astype(np.float32)


def normal(k, shape, layout=PARTITIONABLE):
  shape = tuple(int(s) for s in np.atleast_1d(shape)) if np.ndim(shape) else (int(shape),)
  n = int(np.prod(shape)) if len(shape) else 1
  return normal_from_bits(bits(k, n, layout)).reshape(shape)


def randint_bit(k, n, layout=PARTITIONABLE):
  """jax.random.randint(key, [n], 0, 2): k1,k2=split(key); span 2 makes the
  high-word multiplier (2**16 % 2)**2 % 2 == 0, so the result is bits(k2) & 1
  (nuts.py:551-558 direction draw)."""
  k2 = split(k, 2, layout)[1]
  return (bits(k2, n, layout) & _U32(1)).astype(np.int32)


def randint(k, n, lo, hi, layout=PARTITIONABLE):
  """General jax.random.randint for int32 [lo, hi)."""
  k1, k2 = split(k, 2, layout)
  hb = bits(k1, n, layout).astype(np.uint64)
  lb = bits(k2, n, layout).astype(np.uint64)
  span = np.uint64(int(hi) - int(lo))
  mult = np.uint64((pow(2, 16, int(span)) ** 2) % int(span))
  off = ((hb % span) * mult + (lb % span)) % span
  return (np.int64(lo) + off.astype(np.int64)).astype(np.int32)
