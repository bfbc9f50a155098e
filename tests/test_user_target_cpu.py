"""User-defined targets, host side (no GPU): the run-time build (NVRTC for sm_100a against csrc/pb2_user_target.cuh)
compiles a well-formed source, reports the compiler's log with the user's line numbers for a broken one, and the
Python wrapper validates its arguments."""
import pytest

GAUSS = '''
__device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data) {
  float lp = 0.f;
  for (int d = 0; d < 3; ++d) { g[d] = -x[d]; lp -= 0.5f * x[d] * x[d]; }
  return lp;
}
'''


def test_user_target_compiles_without_a_device(lib_built):
  from probability_b200 import targets
  for dim in (3, 40, 256):   # 1, 2 and 8 state elements per lane
    assert targets.UserTarget(dim, GAUSS).check()


def test_cooperative_user_target_compiles(lib_built):
  from probability_b200 import targets
  src = '''
__device__ float target_log_prob_and_grad(const float* x, float* g, const float* data, int n_data, int lane) {
  float lp = 0.f;
  for (int d = lane; d < 100; d += 32) { g[d] = -x[d]; lp -= 0.5f * x[d] * x[d]; }
  return pb2::warp_sum(lp);
}
'''
  assert targets.UserTarget(100, src, cooperative=True).check()


def test_user_target_compile_error_carries_the_log(lib_built):
  from probability_b200 import _lib, targets
  bad = GAUSS.replace('lp -= 0.5f', 'lp -= undefined_symbol * 0.5f')
  with pytest.raises(_lib.Pb2Error) as e:
    targets.UserTarget(3, bad).check()
  msg = str(e.value)
  assert 'undefined_symbol' in msg and 'user_target.cu(4)' in msg


def test_user_target_argument_checks(lib_built):
  from probability_b200 import targets
  with pytest.raises(TypeError):
    targets.UserTarget(3, lambda x: -0.5 * (x ** 2).sum(-1))
  with pytest.raises(ValueError):
    targets.UserTarget(257, GAUSS)
  with pytest.raises(ValueError):
    targets.UserTarget(3, GAUSS, part_sizes=[1, 1])
  t = targets.UserTarget(3, GAUSS, data=[1.0, 2.0], part_sizes=[1, 2])
  assert t.part_sizes == [1, 2] and t.n_rows == 2
