"""ctypes binding of libpb2.so (the C ABI declared in include/pb2.h).

There is NO fallback: if the shared library is missing, or a compute entry point is
called without a B200 visible, this module raises.  PyTorch is used only to own
device memory and streams; every tensor crosses the boundary as a raw device pointer.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PB2_LIB_PATH') or os.path.join(_HERE, '_C', 'libpb2.so')   # override: A/B builds

PB2_OK = 0
LAYOUT_PARTITIONABLE = 0
LAYOUT_ORIGINAL = 1
LAYOUT_PHILOX = 2
TARGET_EIGHT_SCHOOLS, TARGET_DENSE_GAUSSIAN, TARGET_LOGISTIC, TARGET_STOCH_VOL = 0, 1, 2, 3
TARGET_STOCH_VOL_CONSTRAINED = 4
TARGET_USER = 5
TARGET_STOCH_VOL_CENTERED, TARGET_STOCH_VOL_CENTERED_CONSTRAINED = 6, 7
CSRC_DIR = os.path.join(_HERE, 'csrc')   # the device headers a user-defined target is compiled against (NVRTC)
KERNEL_HMC, KERNEL_NUTS = 0, 1
STEP_SCALAR, STEP_PER_DIM, STEP_PER_CHAIN = 0, 1, 2

c_f32p = C.POINTER(C.c_float)
c_u32p = C.POINTER(C.c_uint32)


class TargetDesc(C.Structure):
  _fields_ = [('kind', C.c_int), ('dim', C.c_int), ('n_rows', C.c_int),
              ('h_a', c_f32p), ('h_b', c_f32p), ('scalar', C.c_float)]


class ChainLayout(C.Structure):
  _fields_ = [('B', C.c_int), ('B_global', C.c_int), ('chain_offset', C.c_int),
              ('rng_layout', C.c_int), ('n_parts', C.c_int), ('part_sizes', C.c_int * 8)]


class RunCfg(C.Structure):
  _fields_ = [('kind', C.c_int), ('num_leapfrog_steps', C.c_int), ('max_tree_depth', C.c_int),
              ('max_energy_diff', C.c_float), ('unrolled_leapfrog_steps', C.c_int),
              ('num_results', C.c_int), ('num_burnin_steps', C.c_int),
              ('num_steps_between_results', C.c_int), ('step_kind', C.c_int),
              ('explicit_step_seeds', C.c_int), ('d_momentum_scale', C.c_void_p),
              ('d_bijector_kind', C.c_void_p), ('d_bijector_low', C.c_void_p), ('d_bijector_high', C.c_void_p)]


TRACE_FIELDS = ['states', 'target_log_prob', 'grads_target_log_prob', 'log_accept_ratio',
                'is_accepted', 'step_size', 'proposed_state', 'proposed_target_log_prob',
                'proposed_grads', 'log_acceptance_correction', 'initial_momentum',
                'final_momentum', 'leapfrogs_taken', 'has_divergence', 'reach_max_depth', 'energy']


class Trace(C.Structure):
  _fields_ = [('d_' + f, C.c_void_p) for f in TRACE_FIELDS]


class DA(C.Structure):
  _fields_ = [('enabled', C.c_int), ('d_state', C.c_void_p), ('reduce_over_ranks', C.c_int)]


COMM_ID_BYTES = 128


class Pb2Error(RuntimeError):
  pass


_lib = None
_lock = threading.Lock()


def load():
  """Load libpb2.so (building is the job of `probability_b200.build` / __graft_entry__.build)."""
  global _lib
  with _lock:
    if _lib is not None:
      return _lib
    if not os.path.exists(LIB_PATH):
      raise Pb2Error(
          'libpb2.so not found at {}: build it with `python -m probability_b200.build` '
          '(probability_b200 has no CPU fallback).'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    vp, i32, f32, ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    sig = {
        'pb2_version': ([], i32),
        'pb2_ctx_create': ([i32, C.POINTER(vp)], i32),
        'pb2_ctx_destroy': ([vp], i32),
        'pb2_ctx_set_stream': ([vp, vp], i32),
        'pb2_ctx_synchronize': ([vp], i32),
        'pb2_last_error': ([vp], C.c_char_p),
        'pb2_launch_count': ([vp], ll),
        'pb2_ctx_set_int': ([vp, C.c_char_p, i32], i32),
        'pb2_target_create': ([vp, C.POINTER(TargetDesc), C.POINTER(vp)], i32),
        'pb2_target_create_user': ([vp, i32, C.c_char_p, i32, c_f32p, ll, C.c_char_p, C.POINTER(vp)], i32),
        'pb2_user_target_check': ([C.c_char_p, i32, i32, C.c_char_p], i32),
        'pb2_target_destroy': ([vp], i32),
        'pb2_target_dim': ([vp], i32),
        'pb2_rng_split': ([c_u32p, i32, i32, c_u32p], i32),
        'pb2_rng_fold_in': ([c_u32p, C.c_uint32, c_u32p], i32),
        'pb2_rng_bits': ([vp, c_u32p, ll, i32, vp], i32),
        'pb2_rng_uniform': ([vp, c_u32p, ll, f32, f32, i32, vp], i32),
        'pb2_rng_normal': ([vp, c_u32p, ll, i32, vp], i32),
        'pb2_rng_randint': ([vp, c_u32p, ll, i32, i32, i32, vp], i32),
        'pb2_logp_grad': ([vp, vp, i32, vp, vp, vp], i32),
        'pb2_logp_grad_transformed': ([vp, vp, i32, vp, vp, vp, vp, vp, vp], i32),
        'pb2_dense_logp_grad_tc': ([vp, vp, i32, vp, vp, vp], i32),
        'pb2_logistic_logp_grad_tc': ([vp, vp, i32, vp, vp, vp], i32),
        'pb2_leapfrog': ([vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp], i32),
        'pb2_run': ([vp, vp, C.POINTER(ChainLayout), C.POINTER(RunCfg), c_u32p, c_u32p, vp, vp, vp, vp,
                     C.POINTER(DA), C.POINTER(Trace), vp], i32),
        'pb2_da_init': ([vp, f32, i32, f32, f32, f32, f32, f32, i32, f32, f32, vp], i32),
        'pb2_da_partial': ([vp, vp, i32, vp], i32),
        'pb2_da_apply': ([vp, vp, i32, ll, vp, vp], i32),
        'pb2_running_moments_update': ([vp, vp, ll, i32, vp], i32),
        'pb2_ess': ([vp, vp, i32, i32, i32, f32, i32, i32, i32, vp], i32),
        'pb2_rhat': ([vp, vp, i32, i32, i32, i32, vp], i32),
        'pb2_rowshard_logistic_grad': ([vp, vp, vp, i32, i32, i32, vp, i32, vp], i32),
        'pb2_rowshard_tc_planes_bytes': ([i32], ll),
        'pb2_rowshard_tc_prepare': ([vp, vp, i32, i32, i32, vp], i32),
        'pb2_rowshard_logistic_grad_tc': ([vp, vp, vp, i32, i32, vp, i32, vp], i32),
        'pb2_rowshard_logistic_finish': ([vp, vp, vp, i32, i32, vp, vp], i32),
        'pb2_hmc_mh_finish': ([vp, i32, i32, i32, i32, i32, c_u32p] + [vp] * 20, i32),
        'pb2_logistic_tc_leapfrog': ([vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp], i32),
        'pb2_lockstep_leapfrog': ([vp, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp], i32),
        'pb2_comm_unique_id': ([vp], i32),
        'pb2_comm_init': ([vp, i32, i32, vp], i32),
        'pb2_comm_destroy': ([vp], i32),
        'pb2_comm_size': ([vp], i32),
        'pb2_comm_rank': ([vp], i32),
        'pb2_comm_allreduce_sum': ([vp, vp, ll], i32),
        'pb2_rowshard_leapfrog': ([vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp],
                                  i32),
    }
    for name, (argtypes, restype) in sig.items():
      try:
        fn = getattr(lib, name)
      except AttributeError:
        continue
      fn.argtypes = argtypes
      fn.restype = restype
    _lib = lib
    return lib


def exported_symbols():
  """Names declared in include/pb2.h (parsed) -- used by the CPU test-suite."""
  import re
  hdr = os.path.join(_HERE, '..', 'include', 'pb2.h')
  txt = open(hdr).read()
  return sorted(set(re.findall(r'\b(pb2_[a-z0-9_]+)\s*\(', txt)))


def check(rc, ctx=None):
  if rc != PB2_OK:
    lib = load()
    msg = lib.pb2_last_error(ctx)
    raise Pb2Error('libpb2 error {}: {}'.format(rc, msg.decode() if msg else '?'))


def key_array(key):
  k = np.ascontiguousarray(np.asarray(key, dtype=np.uint32).reshape(2))
  return k


def u32p(arr):
  return arr.ctypes.data_as(c_u32p)


class Context:
  """One pb2_ctx per (process, device); work is enqueued on torch's current stream."""
  _instances = {}

  def __init__(self, device_index):
    import torch
    if not torch.cuda.is_available():
      raise Pb2Error('probability_b200 needs a CUDA device (B200, sm_100a); none is visible '
                     'and there is no CPU fallback.')
    self.lib = load()
    self.device_index = device_index
    h = C.c_void_p()
    with torch.cuda.device(device_index):
      torch.cuda.current_stream()  # make sure the primary context exists
      check(self.lib.pb2_ctx_create(device_index, C.byref(h)))
    self.handle = h

  @classmethod
  def get(cls, device=None):
    import torch
    if device is None:
      idx = torch.cuda.current_device() if torch.cuda.is_available() else 0
    else:
      device = torch.device(device)
      idx = device.index if device.index is not None else (
          torch.cuda.current_device() if torch.cuda.is_available() else 0)
    if idx not in cls._instances:
      cls._instances[idx] = Context(idx)
    return cls._instances[idx]

  def bind_stream(self):
    import torch
    s = torch.cuda.current_stream(self.device_index).cuda_stream
    check(self.lib.pb2_ctx_set_stream(self.handle, C.c_void_p(s)), self.handle)

  def launch_count(self):
    return int(self.lib.pb2_launch_count(self.handle))

  def set_int(self, name, value):
    check(self.lib.pb2_ctx_set_int(self.handle, name.encode(), int(value)), self.handle)

  def synchronize(self):
    check(self.lib.pb2_ctx_synchronize(self.handle), self.handle)


def ptr(t):
  """Raw device pointer of a contiguous torch tensor (None -> NULL)."""
  if t is None:
    return C.c_void_p(0)
  assert t.is_contiguous(), 'tensors crossing the C ABI must be contiguous'
  return C.c_void_p(t.data_ptr())
