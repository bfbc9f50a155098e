"""GPU parity for the lock-step ROW-SHARDED logistic path (BASELINE config 5 shape, small sizes):
sum over shards == unsharded == oracle, and HMC transitions through it match the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import mcmc as omcmc  # noqa: E402
from oracle import rng as orng  # noqa: E402
from oracle import targets as otargets  # noqa: E402


@pytest.fixture(scope='module')
def tfp():
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  import probability_b200 as tfp_
  return tfp_


def dev():
  return torch.device('cuda', 0)


@pytest.mark.parametrize('n,d,B', [(1000, 24, 130), (5000, 99, 257), (333, 40, 64), (31, 7, 5)])
def test_rowshard_logp_grad_matches_oracle(tfp, n, d, B):
  X, y = tfp.targets.synthetic_logistic_data(n, d, seed=3)
  tg = tfp.targets.RowShardedLogisticRegression(X, y)
  Xb = np.concatenate([X, np.ones((n, 1), np.float32)], 1)
  o64 = otargets.LogisticRegression(Xb.astype(np.float64), y.astype(np.float64), dtype=np.float64)
  th = (0.2 * np.random.default_rng(0).standard_normal((B, d + 1))).astype(np.float32)
  lp, g = tg.log_prob_and_grad(torch.tensor(th, device=dev()))
  lp64, g64 = o64.logp_grad(th.astype(np.float64))
  np.testing.assert_allclose(lp.cpu().numpy(), lp64, rtol=2e-5)
  scale = np.abs(g64).max(axis=1, keepdims=True)
  assert np.max(np.abs(g.cpu().numpy() - g64) / scale) < 2e-5


def test_rowshard_sum_of_shards_equals_unsharded(tfp):
  """What the per-leapfrog all-reduce adds up: packed(shard 0) + packed(shard 1) == packed(all rows)."""
  from probability_b200 import _lib
  n, d, B = 2000, 24, 96
  X, y = tfp.targets.synthetic_logistic_data(n, d, seed=4)
  th = torch.tensor((0.2 * np.random.default_rng(1).standard_normal((B, d + 1))).astype(np.float32), device=dev())
  ctx = _lib.Context.get(dev()); ctx.bind_stream()

  def packed(tg):
    Xd, yd = tg._device_data(dev())
    out = torch.empty(B, tg.dim + 1, device=dev())
    _lib.check(ctx.lib.pb2_rowshard_logistic_grad(ctx.handle, _lib.ptr(Xd), _lib.ptr(yd), tg.n_rows, tg.dim,
                                                  tg.padded_dim, _lib.ptr(th), B, _lib.ptr(out)), ctx.handle)
    return out
  full = packed(tfp.targets.RowShardedLogisticRegression(X, y))
  a = packed(tfp.targets.RowShardedLogisticRegression(X[:1200], y[:1200]))
  b = packed(tfp.targets.RowShardedLogisticRegression(X[1200:], y[1200:]))
  np.testing.assert_allclose((a + b).cpu().numpy(), full.cpu().numpy(), rtol=2e-5, atol=2e-4)
  # deterministic: same launch twice gives identical bits
  np.testing.assert_array_equal(packed(tfp.targets.RowShardedLogisticRegression(X, y)).cpu().numpy(),
                                full.cpu().numpy())


def test_rowshard_hmc_matches_oracle_and_smem_path(tfp):
  n, d, B = 1000, 24, 64
  X, y = tfp.targets.synthetic_logistic_data(n, d, seed=5)
  big = tfp.targets.RowShardedLogisticRegression(X, y)
  small = tfp.targets.LogisticRegression(X, y)                     # warp-per-chain shared-memory kernel
  o32 = otargets.LogisticRegression(small.features_with_bias, y)
  x0 = (0.1 * np.random.default_rng(2).standard_normal((B, d + 1))).astype(np.float32)
  st = torch.tensor(x0, device=dev())
  seed = orng.key(7)
  k_big = tfp.mcmc.HamiltonianMonteCarlo(big, step_size=0.02, num_leapfrog_steps=5)
  k_small = tfp.mcmc.HamiltonianMonteCarlo(small, step_size=0.02, num_leapfrog_steps=5)
  s1, r1 = k_big.one_step(st, k_big.bootstrap_results(st), seed=seed)
  s2, r2 = k_small.one_step(st, k_small.bootstrap_results(st), seed=seed)
  lp0, g0 = o32.logp_grad(x0)
  ref = omcmc.hmc_one_step(o32, x0, lp0, g0, 0.02, 5, seed)
  for s, r in ((s1, r1), (s2, r2)):
    acc = r.is_accepted.cpu().numpy()
    agree = acc == ref['is_accepted']
    assert agree.mean() > 0.95
    np.testing.assert_allclose(s.cpu().numpy()[agree], ref['state'][agree], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(r.log_accept_ratio.cpu().numpy(), ref['log_accept_ratio'], rtol=1e-2, atol=5e-3)
  # a short chain through sample_chain (step loop drives the lock-step leapfrogs)
  states, acc = tfp.mcmc.sample_chain(5, st, kernel=k_big, trace_fn=lambda _, kr: kr.is_accepted, seed=3)
  assert states.shape == (5, B, d + 1) and acc.shape == (5, B)
  assert acc.float().mean() > 0.5


@pytest.mark.gpu
@pytest.mark.parametrize('N,D,B', [(5000, 100, 300), (333, 37, 128), (40, 100, 130)])
def test_rowshard_tensor_core_gradient_matches_fp32_and_float64(N, D, B):
  """pb2_rowshard_logistic_grad_tc (tcgen05, 3xTF32, row segments over the SMs) == the FP32 kernel within rounding,
  and as close to float64 as it."""
  import probability_b200 as tfp
  from probability_b200 import _lib
  dev = torch.device('cuda', 0)
  rng = np.random.default_rng(N + D)
  X = rng.standard_normal((N, D - 1)).astype(np.float32)
  y = (rng.random(N) < 0.5).astype(np.float32)
  tg = tfp.targets.RowShardedLogisticRegression(X, y)
  ctx = _lib.Context.get(dev); ctx.bind_stream()
  Xd, yd = tg._device_data(dev)
  th = (0.2 * rng.standard_normal((B, D))).astype(np.float32)
  tht = torch.tensor(th, device=dev)
  out = {}
  packed = torch.empty(B, D + 1, device=dev)
  _lib.check(ctx.lib.pb2_rowshard_logistic_grad(ctx.handle, _lib.ptr(Xd), _lib.ptr(yd), N, D, tg.padded_dim, _lib.ptr(tht),
                                                B, _lib.ptr(packed)), ctx.handle)
  out['pb2_rowshard_logistic_grad'] = packed.cpu().numpy().astype(np.float64)
  packed = torch.empty(B, D + 1, device=dev)
  _lib.check(ctx.lib.pb2_rowshard_logistic_grad_tc(ctx.handle, _lib.ptr(tg._tc_planes(ctx, dev)), _lib.ptr(yd), N, D,
                                                   _lib.ptr(tht), B, _lib.ptr(packed)), ctx.handle)
  out['pb2_rowshard_logistic_grad_tc'] = packed.cpu().numpy().astype(np.float64)
  X64 = np.concatenate([X, np.ones((N, 1), np.float32)], 1).astype(np.float64)
  z = th.astype(np.float64) @ X64.T
  ll = (y[None] * z - np.logaddexp(0, z)).sum(1)
  g = (y[None] - 1 / (1 + np.exp(-z))) @ X64
  ref = np.concatenate([g, ll[:, None]], 1)
  scale = np.abs(ref).max(1, keepdims=True)
  e_fp = np.max(np.abs(out['pb2_rowshard_logistic_grad'] - ref) / scale)
  e_tc = np.max(np.abs(out['pb2_rowshard_logistic_grad_tc'] - ref) / scale)
  assert e_tc < 3 * max(e_fp, 2e-6), (e_tc, e_fp)
