"""Stateless-seed contract of tfp.random over the JAX key scheme.

Mirrors tfp/internal/samplers.py: `sanitize_seed` (:79-172), `fold_in` (:200-214),
`split_seed` (:217-256), `zeros_seed` (:68-72), plus the draw functions `normal`
(:308-325) and `uniform` (:356-368).  Keys are uint32[2] NumPy arrays (what
`jax.random.key_data` holds); key algebra runs on the host through libpb2's pure-C
entry points, bulk draws run on the GPU (threefry2x32, bit-exact uint32 streams).
"""
import ctypes as C
import hashlib

import numpy as np

from probability_b200 import _lib

PARTITIONABLE = _lib.LAYOUT_PARTITIONABLE
ORIGINAL = _lib.LAYOUT_ORIGINAL
PHILOX = _lib.LAYOUT_PHILOX

# jax_threefry_partitionable: True is the default of current JAX releases.
_default_layout = PARTITIONABLE


def set_threefry_partitionable(flag):
  global _default_layout
  _default_layout = PARTITIONABLE if flag else ORIGINAL


def set_generator(name):
  """Bit generator of every draw and key split: 'threefry' (jax.random's threefry2x32, partitionable layout -- the
  default, bit-exact with the JAX substrate), 'threefry_original' (jax_threefry_partitionable=False) or 'philox'
  (Philox4x32-10, the generator behind tf.random.stateless_* on the reference's TF substrate)."""
  global _default_layout
  layouts = {'threefry': PARTITIONABLE, 'threefry_original': ORIGINAL, 'philox': PHILOX}
  if name not in layouts:
    raise ValueError('generator must be one of {}, got {!r}'.format(sorted(layouts), name))
  _default_layout = layouts[name]


def default_layout():
  return _default_layout


def key(seed):
  """jax.random.key / PRNGKey(int) -> uint32[2] = [hi, lo]."""
  seed = int(seed)
  return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], np.uint32)


def zeros_seed():
  return np.zeros([2], np.uint32)


def _as_key(seed):
  if seed is None:
    raise ValueError('A stateless seed (uint32[2] key or int) is required: this engine follows the '
                     'JAX substrate, where `seed=None` is an error.')
  if isinstance(seed, (int, np.integer)):
    return key(seed)
  k = np.asarray(seed)
  if k.shape != (2,):
    raise ValueError('seed must be an int or a uint32[2] key, got shape {}'.format(k.shape))
  return k.astype(np.uint32)


def fold_in(seed, salt):
  k = np.ascontiguousarray(_as_key(seed))
  out = np.zeros([2], np.uint32)
  _lib.check(_lib.load().pb2_rng_fold_in(_lib.u32p(k), C.c_uint32(int(salt) & 0xFFFFFFFF), _lib.u32p(out)))
  return out


def sanitize_seed(seed, salt=None, name=None):
  del name
  k = _as_key(seed)
  if salt is not None:
    if not isinstance(salt, str):
      raise TypeError('`salt` must be a python `str`, got {}'.format(repr(salt)))
    s = int(hashlib.sha512(str(salt).encode('utf-8')).hexdigest(), 16) % (2**31 - 1)
    k = fold_in(k, s)
  return k


def split_seed(seed, n=2, salt=None, name=None, layout=None):
  del name
  if not isinstance(n, (int, np.integer)):
    raise TypeError('`n` must be a python `int`, got {}'.format(repr(n)))
  k = np.ascontiguousarray(sanitize_seed(seed, salt=salt))
  out = np.zeros([int(n), 2], np.uint32)
  lay = _default_layout if layout is None else layout
  _lib.check(_lib.load().pb2_rng_split(_lib.u32p(k), int(n), lay, _lib.u32p(out)))
  return out


def _draw(fn_name, seed, shape, device, dtype, extra, layout):
  import torch
  shape = tuple(int(s) for s in (shape if hasattr(shape, '__len__') else [shape]))
  n = int(np.prod(shape)) if len(shape) else 1
  ctx = _lib.Context.get(device)
  ctx.bind_stream()
  out = torch.empty(shape, dtype=dtype, device=torch.device('cuda', ctx.device_index))
  k = np.ascontiguousarray(_as_key(seed))
  lay = _default_layout if layout is None else layout
  fn = getattr(ctx.lib, fn_name)
  if fn_name == 'pb2_rng_uniform':
    rc = fn(ctx.handle, _lib.u32p(k), n, extra[0], extra[1], lay, _lib.ptr(out))
  elif fn_name == 'pb2_rng_randint':
    rc = fn(ctx.handle, _lib.u32p(k), n, extra[0], extra[1], lay, _lib.ptr(out))
  else:
    rc = fn(ctx.handle, _lib.u32p(k), n, lay, _lib.ptr(out))
  _lib.check(rc, ctx.handle)
  return out


def bits(seed, shape, device=None, layout=None):
  """uint32 random bits (returned as int32 storage viewed uint32 via .view(torch.uint32))."""
  import torch
  return _draw('pb2_rng_bits', seed, shape, device, torch.int32, None, layout).view(torch.uint32)


def normal(shape, mean=0.0, stddev=1.0, seed=None, device=None, layout=None):
  import torch
  z = _draw('pb2_rng_normal', seed, shape, device, torch.float32, None, layout)
  if stddev != 1.0 or mean != 0.0:
    z = z * stddev + mean
  return z


def uniform(shape, minval=0.0, maxval=None, dtype=None, seed=None, device=None, layout=None):
  import torch
  if dtype is not None and not torch.empty(0, dtype=dtype).is_floating_point():
    if maxval is None:
      raise ValueError('Must specify maxval for integer dtype {}.'.format(dtype))
    return _draw('pb2_rng_randint', seed, shape, device, torch.int32, (int(minval), int(maxval)), layout)
  maxval = 1.0 if maxval is None else maxval
  return _draw('pb2_rng_uniform', seed, shape, device, torch.float32, (float(minval), float(maxval)), layout)
