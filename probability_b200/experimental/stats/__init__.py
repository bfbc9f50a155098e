"""Streaming statistics (tfp/experimental/stats/sample_stats.py:64-460 RunningCovariance / RunningVariance /
RunningMean): running count / mean / sum of squared deviations kept on the device and updated by
pb2_running_moments_update (Welford within a row segment, Chan's merge across segments, fixed order)."""
import numpy as np

from probability_b200 import _lib
from probability_b200.mcmc import _engine


def merge_moment_states(states):
  """Chan's merge, in list order, of `[count, mean[D], sum of squared deviations[D]]` states (float64 inside): the
  moments of the union of the observations.  With the ranks' states gathered in rank order every rank gets the same
  bits (the pooled estimate windowed adaptation uses when the chains are sharded, windowed_sampling.py:322-347 with
  `experimental_chain_axis_names`)."""
  import torch
  D = (states[0].numel() - 1) // 2
  first = states[0].double()
  n, mean, m2 = first[0], first[1:1 + D], first[1 + D:]
  for s in states[1:]:
    s = s.double()
    nb, mb, m2b = s[0], s[1:1 + D], s[1 + D:]
    tot = n + nb
    safe = torch.clamp(tot, min=1.0)
    delta = mb - mean
    mean = mean + delta * (nb / safe)
    m2 = m2 + m2b + delta * delta * (n * nb / safe)
    n = tot
  return torch.cat([n.reshape(1), mean, m2]).float().contiguous()


class RunningVariance(object):
  """`update(new_sample)` folds a batch of observations into the running moments; every row of the flattened
  `[n, D]` batch is one observation (the reference's `update(x, axis=0)`).  Immutable like the reference's: `update`
  returns a new object.  `shapes` are the event shapes of the state parts (the variance has that structure)."""

  def __init__(self, state, shapes, was_list):
    self.state = state          # float32 CUDA tensor [1 + 2 D]: count, mean[D], sum of squared deviations[D]
    self.shapes = list(shapes)
    self.was_list = was_list

  @property
  def dim(self):
    return (self.state.numel() - 1) // 2

  @classmethod
  def from_shape(cls, shapes, device, was_list=True):
    """Zero observations (`RunningVariance.from_shape`)."""
    import torch
    D = sum(_engine.part_sizes_of(shapes))
    return cls(torch.zeros(1 + 2 * D, dtype=torch.float32, device=device), shapes, was_list)

  @classmethod
  def from_stats(cls, num_samples, mean, variance):
    """`RunningVariance.from_stats(num_samples, mean, variance)` (sample_stats.py:182-206): mean / variance are
    tensors or lists of tensors with the event shapes of the state parts."""
    import torch
    was_list = _engine.is_list_like(mean)
    means = list(mean) if was_list else [mean]
    varis = list(variance) if _engine.is_list_like(variance) else [variance]
    dev = None
    for t in means + varis:
      if torch.is_tensor(t) and t.is_cuda:
        dev = t.device
    if dev is None:
      dev = torch.device('cuda', torch.cuda.current_device())
    f = lambda v: torch.as_tensor(np.asarray(v.cpu()) if torch.is_tensor(v) else np.asarray(v, np.float32),
                                  dtype=torch.float32, device=dev)
    means = [f(m) for m in means]
    varis = [f(v) for v in varis]
    n = float(num_samples)
    shapes = [tuple(m.shape) for m in means]
    st = torch.cat([torch.tensor([n], dtype=torch.float32, device=dev)] + [m.reshape(-1) for m in means] +
                   [v.reshape(-1) * n for v in varis])
    return cls(st.contiguous(), shapes, was_list)

  def update(self, new_sample, axis=0):
    """new_sample: tensor / list of tensors `[n, *event]` (or `[a, b, *event]`, flattened over the leading axes)."""
    import torch
    del axis
    parts = list(new_sample) if _engine.is_list_like(new_sample) else [new_sample]
    sizes = _engine.part_sizes_of(self.shapes)
    flat = []
    for p, n in zip(parts, sizes):
      p = torch.as_tensor(p, dtype=torch.float32, device=self.state.device)
      flat.append(p.reshape(-1, n))
    x = (flat[0] if len(flat) == 1 else torch.cat(flat, dim=1)).contiguous()
    st = self.state.clone()
    ctx = _lib.Context.get(st.device)
    ctx.bind_stream()
    _lib.check(ctx.lib.pb2_running_moments_update(ctx.handle, _lib.ptr(x), x.shape[0], x.shape[1], _lib.ptr(st)),
               ctx.handle)
    return RunningVariance(st, self.shapes, self.was_list)

  def merged_over_ranks(self, dist, group=None):
    """The moments of all ranks' observations (every rank gets the same object state)."""
    import torch
    gathered = [torch.empty_like(self.state) for _ in range(dist.get_world_size(group))]
    dist.all_gather(gathered, self.state.contiguous(), group=group)
    return RunningVariance(merge_moment_states(gathered), self.shapes, self.was_list)

  @property
  def num_samples(self):
    return self.state[0]

  def _unflat(self, v):
    outs, off = [], 0
    for s, n in zip(self.shapes, _engine.part_sizes_of(self.shapes)):
      outs.append(v[off:off + n].reshape(s))
      off += n
    return outs if self.was_list else outs[0]

  @property
  def mean(self):
    return self._unflat(self.state[1:1 + self.dim])

  def variance_flat(self, ddof=0):
    """[D] population (ddof = 0, the reference default) variance; ones while there are no observations."""
    import torch
    n = self.state[0] - float(ddof)
    v = self.state[1 + self.dim:] / torch.clamp(n, min=1.0)
    return torch.where(self.state[0] > float(ddof), v, torch.ones_like(v))

  def variance(self, ddof=0):
    return self._unflat(self.variance_flat(ddof))
