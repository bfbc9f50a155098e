"""Multi-GPU plumbing of the hot path: one process per GPU, one NCCL communicator per libpb2 context.

The reference expresses cross-device work through named axes (tfp/internal/distribute_lib.py:147-242: psum,
reduce_logsumexp, pbroadcast; experimental/mcmc/sharded.py).  Here the collectives that sit ON the path are issued by
the library itself (include/pb2.h "multi-GPU group"):

  * chain-sharded runs: the dual-averaging accept statistic of every adapting transition (pb2_run,
    `experimental_reduce_chain_axis_names`);
  * row-sharded data: the per-leapfrog gradient all-reduce inside pb2_rowshard_leapfrog.

`init_comm()` attaches the communicator: rank 0 creates the NCCL id (pb2_comm_unique_id), torch.distributed only
carries the 128 id bytes to the other ranks (any backend: nccl or gloo), every rank calls pb2_comm_init.
"""
import ctypes as C

import numpy as np

from probability_b200 import _lib


def init_comm(device=None, process_group=None):
  """Attach an NCCL communicator spanning `process_group` (default: the world) to this process's context.
  Returns the number of ranks.  A no-op returning 1 when torch.distributed is not initialised or has one rank."""
  import torch
  import torch.distributed as dist
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) < 2:
    return 1
  ctx = _lib.Context.get(device)
  n = int(ctx.lib.pb2_comm_size(ctx.handle))
  world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
  if n == world:
    return n
  if n > 1:
    raise _lib.Pb2Error('this context already has a communicator of {} ranks'.format(n))
  buf = (C.c_ubyte * _lib.COMM_ID_BYTES)()
  if rank == 0:
    _lib.check(ctx.lib.pb2_comm_unique_id(buf))
  backend = dist.get_backend(process_group)
  dev = torch.device('cuda', ctx.device_index) if 'nccl' in str(backend) else torch.device('cpu')
  t = torch.tensor(np.frombuffer(buf, dtype=np.uint8).copy(), device=dev)
  src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
  dist.broadcast(t, src=src, group=process_group)
  idb = t.cpu().numpy().tobytes()
  with torch.cuda.device(ctx.device_index):
    _lib.check(ctx.lib.pb2_comm_init(ctx.handle, world, rank, C.c_char_p(idb)), ctx.handle)
  return world


def comm_size(device=None):
  """Ranks of the communicator attached to this process's context (1: none)."""
  ctx = _lib.Context.get(device)
  return int(ctx.lib.pb2_comm_size(ctx.handle))


def destroy_comm(device=None):
  ctx = _lib.Context.get(device)
  _lib.check(ctx.lib.pb2_comm_destroy(ctx.handle), ctx.handle)


def all_reduce_sum(t):
  """In-place sum of a float32 CUDA tensor over the ranks of the attached communicator (pb2_comm_allreduce_sum)."""
  ctx = _lib.Context.get(t.device)
  ctx.bind_stream()
  _lib.check(ctx.lib.pb2_comm_allreduce_sum(ctx.handle, _lib.ptr(t), t.numel()), ctx.handle)
  return t


# ---- named axes (distribute_lib.py:33-73 canonicalize_named_axis, get_axis_index / get_axis_size) ---------------
_AXES = {}   # name -> (index, size) or a torch.distributed process group


def canonicalize_named_axis(named_axes):
  """distribute_lib.canonicalize_named_axis: None -> [], 'a' -> ['a'], iterables -> list."""
  if named_axes is None:
    return []
  if isinstance(named_axes, str):
    return [named_axes]
  return list(named_axes)


def register_axis(name, index=None, size=None, process_group=None):
  """Bind a named axis to a position: either an explicit (index, size) -- e.g. several shards driven by one process --
  or a torch.distributed process group (the member index is this process's rank in it)."""
  if process_group is not None:
    _AXES[name] = process_group
  else:
    if index is None or size is None or not 0 <= int(index) < int(size):
      raise ValueError('register_axis needs 0 <= index < size or a process group')
    _AXES[name] = (int(index), int(size))


def unregister_axis(name):
  _AXES.pop(name, None)


def _axis(name):
  import torch.distributed as dist
  a = _AXES.get(name)
  if isinstance(a, tuple):
    return a
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(a), dist.get_world_size(a)
  return 0, 1


def get_axis_index(name):
  return _axis(name)[0]


def get_axis_size(name):
  return _axis(name)[1]


def fold_in_axis_index(seed, axis_name=None):
  """distribute_lib.fold_in_axis_index (:193-207): fold the index of this process along every named axis into the key,
  i.e. a different key on every member (independent chains per shard, experimental/mcmc/sharded.py:63-74).  An int is
  taken as the axis index itself.  Chain sharding through `experimental_chain_shard` does NOT need this: there the RNG
  counters are the global chain indices and the sharded run equals the unsharded one bit for bit."""
  from probability_b200 import random as pb_random
  if axis_name is None:
    return seed
  k = pb_random.sanitize_seed(seed)
  if isinstance(axis_name, int):
    return pb_random.fold_in(k, int(axis_name))
  for name in canonicalize_named_axis(axis_name):
    k = pb_random.fold_in(k, get_axis_index(name))
  return k
