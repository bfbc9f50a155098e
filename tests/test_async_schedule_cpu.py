"""CPU proof-by-enumeration of the index arithmetic of the asynchronous-lane tile NUTS kernel
(probability_b200/csrc/pb2_tile_nuts.cu) against the reference's checkpoint tables.

The reference (tfp/mcmc/nuts.py:1013-1071) writes, at even leaf i of a doubling, the pair (momentum, rho) into slot
popcount(i) and, at odd leaf i with t trailing ones, checks the U-turn criterion against slots
[popcount(i) - t, popcount(i)), i.e. against the checkpoints of leaves e_j = i - 2^j + 1, j = 1..t (the first leaves
of the 2-, 4-, .., 2^t-leaf subtrees that close at leaf i).  The kernel stores the same checkpoints in four places
and runs every lane on a shared 32-tick chunk clock:
  * `ckl`   : the previous even leaf's checkpoint (shared memory)                     -> the j = 1 check
  * local slot popcount(tick): written only at ticks with tick % 4 == 0 ("keep")      -> the j = 2..5 checks
  * hi slot popcount(chunk index): first leaf of a chunk, stored iff the chunk index
              is even and it is not the doubling's last chunk                         -> the j >= 6 checks (tick 31)
  * START lanes run doubling d = 0..4 in the aligned block [2^d, 2^(d+1)) of ONE chunk (d = 0: tick 1), masking
    subtrees larger than their doubling (jmax = floor(log2 tick)).
This test replays that storage discipline for every doubling up to depth 10 and asserts that every check the
reference performs finds exactly the reference's checkpoint leaf, and that no other check is performed."""
import numpy as np

from oracle import mcmc as omcmc

CHUNK = 32
S0 = 5


def _trailing_ones(i):
  t = 0
  while i & 1:
    t += 1
    i >>= 1
  return t


def _popc(i):
  return bin(i).count('1')


def _reference_checks(depth):
  """{leaf i: [checkpoint leaves checked at i]} for one doubling of 2^depth leaves, from the reference tables."""
  write, read = omcmc.write_read_instructions(depth) if depth > 0 else (np.array([0]), np.zeros((1, 2), int))
  n = 1 << depth
  slot_owner = {}
  out = {}
  for i in range(n):
    if i % 2 == 0:
      slot_owner[int(write[i])] = i
    lo, hi = int(read[i][0]), int(read[i][1])
    out[i] = sorted(slot_owner[s] for s in range(lo, hi))
  return out


def _kernel_chunk_doubling(it):
  """Replay a CHUNK-state lane through doubling `it` >= 5; returns {leaf: [checkpoint leaves it checked]}."""
  nchunks = 1 << (it - S0)
  hi_slots = {}
  out = {}
  for ihi in range(nchunks):
    local, ckl = {}, None
    hi_slot_w = _popc(ihi) if (ihi % 2 == 0 and ihi + 1 < nchunks) else -1
    t_hi = _trailing_ones(ihi)
    for i in range(CHUNK):
      leaf = CHUNK * ihi + i
      pc, ones, odd = _popc(i), _trailing_ones(i), i & 1
      checked = []
      if not odd:
        ckl = leaf
        if i % 4 == 0:
          local[pc] = leaf
        if i == 0 and hi_slot_w >= 0:
          hi_slots[hi_slot_w] = leaf
      else:
        jmax = S0
        if jmax >= 1:
          checked.append(ckl)                      # j = 1
        for k in range(pc - ones, pc - 1):         # j = pc - k = ones .. 2
          if pc - k <= jmax:
            checked.append(local[k])
        if i == CHUNK - 1:
          for jj in range(1, t_hi + 1):            # j = 5 + jj
            checked.append(hi_slots[_popc(ihi - (1 << jj) + 1)])
      out[leaf] = sorted(checked)
  return out


def _kernel_start_doubling(d):
  """Replay a START-state lane through doubling d <= 4, which sits in ticks [2^d, 2^(d+1)) (d = 0: tick 1)."""
  lo = 1 << d if d >= 1 else 1
  local, ckl = {}, None
  out = {}
  for tick in range(lo, lo + (1 << d)):
    leaf = tick - lo
    pc, ones, odd = _popc(tick), _trailing_ones(tick), tick & 1
    jmax = tick.bit_length() - 1                   # 31 - clz(tick)
    checked = []
    if not odd:
      ckl = leaf
      if tick % 4 == 0:
        local[pc] = leaf
    else:
      if jmax >= 1:
        checked.append(ckl)
      for k in range(pc - ones, pc - 1):
        if pc - k <= jmax:
          checked.append(local[k])
    out[leaf] = sorted(checked)
  return out


def test_chunk_lanes_check_exactly_the_reference_checkpoints():
  for it in range(S0, 11):
    ref = _reference_checks(it)
    got = _kernel_chunk_doubling(it)
    assert got == ref, it


def test_start_lanes_check_exactly_the_reference_checkpoints():
  for d in range(0, S0):
    ref = _reference_checks(d)
    got = _kernel_start_doubling(d)
    assert got == ref, d


def test_start_blocks_tile_one_chunk_and_multinomial_key_indices():
  """doublings 0..4 occupy ticks 1, [2,4), [4,8), [8,16), [16,32): every tick but 0 exactly once; the multinomial key
  of leaf l of doubling d has index (2^d - 1) + l in the transition's key row (nuts.py:622-625) = tick - 1."""
  seen = []
  for d in range(S0):
    lo = 1 << d if d >= 1 else 1
    for leaf in range(1 << d):
      tick = lo + leaf
      seen.append(tick)
      assert (1 << d) - 1 + leaf == tick - 1
  assert sorted(seen) == list(range(1, CHUNK))
  boundaries = [t for t in range(1, CHUNK) if (t + 1) & t == 0]     # doubling d ends after tick 2^(d+1) - 1
  assert boundaries == [1, 3, 7, 15, 31]


def test_ticket_fifo_ring_never_aliases():
  """The FIFO ring has B + lanes + 64 slots: live tickets = chains waiting in the queue (<= B) + idle lanes holding a
  ticket beyond the tail (<= lanes) always span fewer slots than the ring, whatever the interleaving."""
  rng = np.random.default_rng(0)
  B, lanes = 37, 64
  cap = B + lanes + 64
  head = 0                      # next ticket
  tail = B                      # next push
  queued = list(range(B))       # tickets of pushed, not yet consumed chains
  waiting = []                  # tickets drawn by idle lanes, not yet served
  running = 0
  for _ in range(20000):
    ev = rng.integers(3)
    if ev == 0 and len(waiting) + running < lanes:          # an idle lane draws a ticket
      waiting.append(head); head += 1
    elif ev == 1 and running > 0:                            # a running lane finishes a transition: push
      queued.append(tail); tail += 1; running -= 1
    elif ev == 2 and waiting:                                # a waiting lane polls its slot
      tk = waiting[0]
      if tk in queued:
        queued.remove(tk); waiting.pop(0); running += 1
    live = set(queued) | set(waiting)     # a pushed chain and the lane waiting for it share ticket and slot
    if live:
      assert max(live) - min(live) < cap
      assert len({t % cap for t in live}) == len(live)
