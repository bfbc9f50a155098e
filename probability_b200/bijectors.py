"""Elementwise event-space bijectors (tfp/bijectors/{identity,exp,softplus,sigmoid}.py): the set the named targets'
`default_event_space_bijector`s are made of (e.g. vectorized_stochastic_volatility.py:346-356).  Inside the CUDA
transition kernels a bijector is a per-dimension code (pb2.h PB2_BIJECTOR_*; pb2_targets.cuh TransformedT evaluates
forward, its derivative and the forward log-det-Jacobian); the torch methods below are the same maps for the host
side of TransformedTransitionKernel (mapping traced states back to the constrained space)."""
import numpy as np

IDENTITY, EXP, SOFTPLUS, SIGMOID = 0, 1, 2, 3


class Bijector(object):
  code = IDENTITY
  low, high = 0.0, 1.0

  def forward(self, x):
    raise NotImplementedError

  def inverse(self, y):
    raise NotImplementedError

  def forward_log_det_jacobian(self, x, event_ndims=0):
    raise NotImplementedError

  def _reduce(self, v, event_ndims):
    return v.sum(tuple(range(v.dim() - event_ndims, v.dim()))) if event_ndims else v


class Identity(Bijector):
  code = IDENTITY

  def forward(self, x):
    return x

  def inverse(self, y):
    return y

  def forward_log_det_jacobian(self, x, event_ndims=0):
    import torch
    return self._reduce(torch.zeros_like(x), event_ndims)


class Exp(Bijector):
  code = EXP

  def forward(self, x):
    return x.exp()

  def inverse(self, y):
    return y.log()

  def forward_log_det_jacobian(self, x, event_ndims=0):
    return self._reduce(x, event_ndims)


class Softplus(Bijector):
  code = SOFTPLUS

  def forward(self, x):
    import torch
    return torch.nn.functional.softplus(x)

  def inverse(self, y):
    import torch
    return y + torch.log(-torch.expm1(-y))          # softplus_inverse (math/generic.py:530-583)

  def forward_log_det_jacobian(self, x, event_ndims=0):
    import torch
    return self._reduce(-torch.nn.functional.softplus(-x), event_ndims)


class Sigmoid(Bijector):
  code = SIGMOID

  def __init__(self, low=0.0, high=1.0):
    self.low, self.high = float(low), float(high)
    if not self.high > self.low:
      raise ValueError('Sigmoid bijector needs high > low')

  def forward(self, x):
    import torch
    return self.low + (self.high - self.low) * torch.sigmoid(x)

  def inverse(self, y):
    import torch
    z = (y - self.low) / (self.high - self.low)
    return torch.log(z) - torch.log1p(-z)

  def forward_log_det_jacobian(self, x, event_ndims=0):
    import torch
    sp = torch.nn.functional.softplus
    return self._reduce(float(np.log(self.high - self.low)) - sp(-x) - sp(x), event_ndims)
