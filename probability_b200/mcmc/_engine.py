"""Host-side glue between the tfp.mcmc-shaped Python classes and libpb2's pb2_run.

State convention: the reference's list of state parts (each `[chains, *event]`) is
flattened to one contiguous float32 CUDA tensor `[B, D]`; `part_sizes` records the
split so that momentum draws use one key per part exactly like hmc.py:684-695.
"""
import collections
import ctypes as C

import numpy as np

from probability_b200 import _lib
from probability_b200 import random as pb_random
from probability_b200 import targets as pb_targets


def is_list_like(x):
  return isinstance(x, (list, tuple)) and not hasattr(x, '_fields')


def flatten_state(state):
  """-> (x [B,D] float32 contiguous, part_shapes, was_list)."""
  import torch
  was_list = is_list_like(state)
  parts = list(state) if was_list else [state]
  parts = [p if torch.is_tensor(p) else torch.as_tensor(np.asarray(p, np.float32)) for p in parts]
  if not parts:
    raise ValueError('empty state')
  dev = None
  for p in parts:
    if p.is_cuda:
      dev = p.device
  if dev is None:
    if not torch.cuda.is_available():
      raise _lib.Pb2Error('probability_b200 needs a CUDA device (B200); none is visible and there is '
                          'no CPU fallback.')
    dev = torch.device('cuda', torch.cuda.current_device())
  nb = None
  shapes = []
  flat = []
  for p in parts:
    p = p.to(device=dev, dtype=torch.float32)
    if p.dim() == 0:
      p = p.reshape(1)
    b = p.shape[0]
    if nb is None:
      nb = b
    elif nb != b:
      raise ValueError('all state parts must share the leading (chain) dimension; got {} and {}'.format(nb, b))
    shapes.append(tuple(p.shape[1:]))
    flat.append(p.reshape(b, -1))
  x = flat[0] if len(flat) == 1 else torch.cat(flat, dim=1)
  return x.contiguous(), shapes, was_list


def part_sizes_of(shapes):
  return [int(np.prod(s)) if len(s) else 1 for s in shapes]


def unflatten(x, shapes, was_list, lead=()):
  """x: [*lead, B, D] -> list of [*lead, B, *shape] (or a single tensor)."""
  sizes = part_sizes_of(shapes)
  outs = []
  off = 0
  for s, n in zip(shapes, sizes):
    piece = x[..., off:off + n]
    outs.append(piece.reshape(tuple(x.shape[:-1]) + tuple(s)))
    off += n
  return outs if was_list else outs[0]


def require_target(fn):
  if not isinstance(fn, pb_targets.Target):
    raise TypeError(
        'target_log_prob_fn must be a probability_b200.targets.Target (the transition kernels are '
        'persistent CUDA kernels with fused log-prob/gradient code; arbitrary Python callables cannot '
        'run inside them and there is no CPU/autodiff fallback). Got: {!r}'.format(fn))
  return fn


def step_size_tensor(step_size, B, D, shapes, device):
  """Map the reference's step-size forms (scalar | per-part list | [D] | [B,1]) to
  (tensor, step_kind)."""
  import torch
  if is_list_like(step_size):
    parts = list(step_size)
    sizes = part_sizes_of(shapes)
    if len(parts) == 1:
      parts = parts * len(sizes)
    if len(parts) != len(sizes):
      raise ValueError('There should be exactly one `step_size` or it should have same length as '
                       '`current_state`.')
    parts = [torch.as_tensor(s, dtype=torch.float32, device=device) for s in parts]
    # per-chain forms ([B] + [1]*event_rank) in every part -> ONE per-chain step (they must agree: the kernels take
    # one step size per chain); anything that broadcasts against the event shape -> per dimension
    per_chain = [s.dim() == len(shp) + 1 and s.shape[0] == B and s.numel() == B for s, shp in zip(parts, shapes)]
    if all(per_chain):
      flat = [s.reshape(B) for s in parts]
      for f in flat[1:]:
        if not torch.equal(f, flat[0]):
          raise ValueError('per-chain step sizes must be the same for every state part (the kernels take one '
                           'step size per chain)')
      return flat[0].contiguous().clone(), _lib.STEP_PER_CHAIN
    cols = []
    for s, n, shp in zip(parts, sizes, shapes):
      if s.dim() > len(shp):
        raise ValueError('unsupported step_size part of shape {} for a state part of event shape {} ({} chains): '
                         'a part must broadcast against the event shape, or every part must be per-chain '
                         '[chains, 1, ...]'.format(tuple(s.shape), tuple(shp), B))
      cols.append(torch.broadcast_to(s, shp if len(shp) else (1,)).reshape(-1))
    out = torch.cat(cols).contiguous()
    if out.numel() != D:
      raise ValueError('per-part step sizes cover {} dimensions but the state has {}'.format(out.numel(), D))
    return out, _lib.STEP_PER_DIM
  s = torch.as_tensor(step_size, dtype=torch.float32, device=device)
  if s.dim() == 0 or s.numel() == 1:
    return s.reshape(1).contiguous().clone(), _lib.STEP_SCALAR
  if s.dim() == 1 and s.shape[0] == D:
    return s.contiguous().clone(), _lib.STEP_PER_DIM
  if s.dim() >= 2 and s.shape[0] == B and s.numel() == B:
    return s.reshape(B).contiguous().clone(), _lib.STEP_PER_CHAIN
  raise ValueError('unsupported step_size shape {} for state [{}, {}]'.format(tuple(s.shape), B, D))


ChainShard = collections.namedtuple('ChainShard', ['chain_offset', 'num_chains_global'])


def run(target, x, lp, g, step, step_kind, shapes, *, kind, num_results, num_burnin_steps=0,
        num_steps_between_results=0, seed=None, step_seeds=None, num_leapfrog_steps=1,
        max_tree_depth=10, max_energy_diff=1000.0, unrolled_leapfrog_steps=1, want=(),
        da_state=None, shard=None, leapfrog_total=None, layout=None, da_over_ranks=False,
        momentum_scale=None):
  """Calls pb2_run.  x, lp, g, step are updated IN PLACE (pass clones to keep inputs).
  Returns (trace dict of tensors with leading R, final pass-along seed, step seeds)."""
  import torch
  B, D = x.shape
  ctx = _lib.Context.get(x.device)
  ctx.bind_stream()
  sizes = part_sizes_of(shapes)
  lay = _lib.ChainLayout()
  lay.B = B
  lay.B_global = B if shard is None else int(shard.num_chains_global)
  lay.chain_offset = 0 if shard is None else int(shard.chain_offset)
  lay.rng_layout = pb_random.default_layout() if layout is None else layout
  lay.n_parts = len(sizes)
  if len(sizes) > 8:
    raise ValueError('at most 8 state parts are supported')
  for i, n in enumerate(sizes):
    lay.part_sizes[i] = n
  cfg = _lib.RunCfg(kind=kind, num_leapfrog_steps=int(num_leapfrog_steps), max_tree_depth=int(max_tree_depth),
                    max_energy_diff=float(max_energy_diff),
                    unrolled_leapfrog_steps=int(unrolled_leapfrog_steps), num_results=int(num_results),
                    num_burnin_steps=int(num_burnin_steps),
                    num_steps_between_results=int(num_steps_between_results), step_kind=int(step_kind),
                    explicit_step_seeds=0 if step_seeds is None else 1,
                    d_momentum_scale=None if momentum_scale is None else momentum_scale.data_ptr())
  bij = target.bijector_arrays(x.device) if hasattr(target, 'bijector_arrays') else None
  if bij is not None:   # TransformedTransitionKernel: x is the unconstrained state
    cfg.d_bijector_kind, cfg.d_bijector_low, cfg.d_bijector_high = (b.data_ptr() for b in bij)
  n_steps = int(num_burnin_steps) + 1 + (int(num_results) - 1) * (1 + int(num_steps_between_results))
  if step_seeds is None:
    h_seed = np.ascontiguousarray(np.asarray(seed, np.uint32).reshape(2)).copy()
    h_steps = np.zeros([n_steps, 2], np.uint32)
  else:
    h_seed = np.zeros([2], np.uint32)
    h_steps = np.ascontiguousarray(np.asarray(step_seeds, np.uint32).reshape(n_steps, 2)).copy()
  R = int(num_results)
  dev = x.device
  spec = {
      'states': ((R, B, D), torch.float32), 'target_log_prob': ((R, B), torch.float32),
      'grads_target_log_prob': ((R, B, D), torch.float32), 'log_accept_ratio': ((R, B), torch.float32),
      'is_accepted': ((R, B), torch.uint8), 'step_size': ((R,), torch.float32),
      'proposed_state': ((R, B, D), torch.float32), 'proposed_target_log_prob': ((R, B), torch.float32),
      'proposed_grads': ((R, B, D), torch.float32), 'log_acceptance_correction': ((R, B), torch.float32),
      'initial_momentum': ((R, B, D), torch.float32), 'final_momentum': ((R, B, D), torch.float32),
      'leapfrogs_taken': ((R, B), torch.int32), 'has_divergence': ((R, B), torch.uint8),
      'reach_max_depth': ((R, B), torch.uint8), 'energy': ((R, B), torch.float32),
  }
  out = {}
  tr = _lib.Trace()
  for name in want:
    shape, dt = spec[name]
    if name == 'step_size' and step_kind != _lib.STEP_SCALAR:
      continue
    t = torch.empty(shape, dtype=dt, device=dev)
    out[name] = t
    setattr(tr, 'd_' + name, t.data_ptr())
  da = _lib.DA(enabled=0 if da_state is None else 1,
               d_state=None if da_state is None else da_state.data_ptr(),
               reduce_over_ranks=1 if da_over_ranks else 0)
  rc = ctx.lib.pb2_run(ctx.handle, target.handle(ctx), C.byref(lay), C.byref(cfg), _lib.u32p(h_seed),
                       _lib.u32p(h_steps), _lib.ptr(x), _lib.ptr(lp), _lib.ptr(g), _lib.ptr(step),
                       C.byref(da), C.byref(tr), _lib.ptr(leapfrog_total))
  _lib.check(rc, ctx.handle)
  for name in ('is_accepted', 'has_divergence', 'reach_max_depth'):
    if name in out:
      out[name] = out[name].bool()
  return out, h_seed, h_steps


def check_shard_axis_names(shard_axis_names):
  """State parts sharded over a named axis need the dot products of the transition psum'd over that axis
  (hmc.py:706-716, nuts.py:994-998).  This engine keeps every chain's whole state on one GPU -- it shards chains
  (ChainShard, Sharded) and data rows (RowShardedLogistic) instead -- so only axes of size 1 are accepted."""
  from probability_b200 import distribute
  if not shard_axis_names:
    return
  def names(s):
    if isinstance(s, str):
      return [s]
    out = []
    for e in s:
      if e is not None:
        out.extend(names(e))
    return out
  for name in names(shard_axis_names):
    if distribute.get_axis_size(name) > 1:
      raise NotImplementedError(
          'experimental_shard_axis_names: state parts sharded over axis {!r} (size {}) are not supported; shard the '
          'chains (ChainShard / Sharded) or the data rows (RowShardedLogistic)'.format(name, distribute.get_axis_size(name)))
