"""Philox4x32-10 (the bit generator north_star names; tf.random.stateless_* on the reference's TF substrate,
tfp/internal/samplers.py:249-250,324-325,367-368): the oracle against the Random123 known-answer vectors
(Random123 kat_vectors, philox4x32 10 rounds), and the library's host-side key split against the oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import rng as orng

KAT = [  # key, counter -> output
    ((0, 0), (0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff), (0xffffffff,) * 4, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize('key,ctr,out', KAT)
def test_oracle_philox_known_answers(key, ctr, out):
  got = orng.philox4x32_10(key[0], key[1], *ctr)
  assert tuple(int(v) for v in got) == out


def test_philox_stream_layout():
  """Element j = word j % 4 of block j // 4; a shorter draw is a prefix of a longer one; split = bits(2n)."""
  k = orng.key(5)
  b = orng.bits(k, 11, orng.PHILOX)
  blk2 = orng.philox4x32_10(k[0], k[1], 2, 0, 0, 0)
  assert [int(v) for v in b[8:11]] == [int(v) for v in blk2[:3]]
  np.testing.assert_array_equal(orng.bits(k, 7, orng.PHILOX), b[:7])
  np.testing.assert_array_equal(orng.split(k, 3, orng.PHILOX), b[:6].reshape(3, 2))
  assert len(np.unique(orng.bits(k, 4096, orng.PHILOX))) > 4090


def test_host_split_matches_oracle(lib_built):
  lib = C.CDLL(lib_built)
  lib.pb2_rng_split.argtypes = [C.POINTER(C.c_uint32), C.c_int, C.c_int, C.POINTER(C.c_uint32)]
  for seed in (0, 17, 2**40 + 3):
    k = np.ascontiguousarray(orng.key(seed))
    for n in (1, 2, 5):
      out = np.zeros([n, 2], np.uint32)
      rc = lib.pb2_rng_split(k.ctypes.data_as(C.POINTER(C.c_uint32)), n, 2, out.ctypes.data_as(C.POINTER(C.c_uint32)))
      assert rc == 0
      np.testing.assert_array_equal(out, orng.split(k, n, orng.PHILOX))
  out = np.zeros([2, 2], np.uint32)
  assert lib.pb2_rng_split(k.ctypes.data_as(C.POINTER(C.c_uint32)), 2, 3, out.ctypes.data_as(C.POINTER(C.c_uint32))) != 0


def test_set_generator_names():
  from probability_b200 import random as pb_random
  try:
    pb_random.set_generator('philox')
    assert pb_random.default_layout() == pb_random.PHILOX == 2
    pb_random.set_generator('threefry_original')
    assert pb_random.default_layout() == pb_random.ORIGINAL
    with pytest.raises(ValueError):
      pb_random.set_generator('mt19937')
  finally:
    pb_random.set_generator('threefry')
  assert pb_random.default_layout() == pb_random.PARTITIONABLE
