"""Oracle MCMC: NumPy restatement of the tfp.mcmc HMC / NUTS path (TEST INFRASTRUCTURE).

All state is one flat float32 array x[B, D]; `part_sizes` says how D splits into
the reference's state parts (one momentum key per part, counters row-major inside
the part's [B, size] shape -- hmc.py:684-695, nuts.py:515-523).

Reference lines followed:
  leapfrog               mcmc/internal/leapfrog_integrator.py:280-309,330-355
  hmc_one_step           mcmc/hmc.py:661-729,780-875 + metropolis_hastings.py:181-254
  nuts tables            mcmc/nuts.py:1013-1071 (pins nuts_test.py:191-215)
  nuts_one_step          mcmc/nuts.py:321-445,512-946 (batched lock-step, literal)
  nuts_one_step_chain    same arithmetic, per-chain early exit
  DualAveraging          mcmc/dual_averaging_step_size_adaptation.py:419-475,543-609
  sample_chain           mcmc/sample.py:311-383
"""
import numpy as np

from oracle import rng as orng

f32 = np.float32
NEG_INF = f32(-np.inf)


# ----------------------------------------------------------------------------
# helpers
def _step_b(step_size, B, D):
  """Broadcast a step size (scalar | [D] | [B,1] | [B]) to [B, D] float32."""
  s = np.asarray(step_size, f32)
  if s.ndim == 1 and s.size == B and B != D:
    s = s[:, None]
  return np.broadcast_to(s, (B, D)).astype(f32)


def softplus(x):
  x = np.asarray(x, f32)
  with np.errstate(over='ignore', invalid='ignore'):
    return (np.log1p(np.exp(-np.abs(x))) + np.maximum(x, f32(0))).astype(f32)


def log_add_exp(x, y):
  """math/generic.py:585-611."""
  with np.errstate(invalid='ignore'):
    larger = np.maximum(x, y)
    return (larger + softplus((x - larger) + (y - larger))).astype(f32)


def draw_momentum(part_keys, B, part_sizes, layout, chain_offset=0, B_global=None):
  """One normal draw of shape [B_global, size_p] per part; rows [offset, offset+B)."""
  Bg = B if B_global is None else B_global
  cols = []
  for k, sz in zip(part_keys, part_sizes):
    full = orng.normal(k, (Bg, sz), layout)
    cols.append(full[chain_offset:chain_offset + B])
  return np.concatenate(cols, axis=1).astype(f32)


def _uniform_b(k, B, layout, chain_offset=0, B_global=None):
  Bg = B if B_global is None else B_global
  return orng.uniform(k, (Bg,), layout=layout)[chain_offset:chain_offset + B]


# ----------------------------------------------------------------------------
def leapfrog(target, m, x, lp, g, eps, L, inv_mass=None):
  """leapfrog_integrator.py:280-309 (+ _one_step :330-355); eps is [B,D] (signed).  `inv_mass` [D]: diagonal of the
  inverse mass matrix -- the position moves along the VELOCITY inv_mass * momentum, the gradient of the kinetic energy
  (kinetic_energy_fn hook :267-275; experimental/mcmc/preconditioned_hmc.py, preconditioning_utils.py)."""
  m = np.asarray(m, f32); x = np.asarray(x, f32); g = np.asarray(g, f32)
  h = (f32(0.5) * eps).astype(f32)
  v = (m + h * g).astype(f32)
  for _ in range(int(L)):
    if inv_mass is None:
      x = (x + eps * v).astype(f32)
    else:
      x = (x + eps * (f32(1) * inv_mass * v).astype(f32)).astype(f32)
    lp, g = target.logp_grad(x)
    v = (v + eps * g).astype(f32)
  m = (v - h * g).astype(f32)
  return m, x, lp, g


def safe_sum(terms):
  """mcmc/internal/util.py:205-235."""
  with np.errstate(invalid='ignore', over='ignore'):
    s = terms[0]
    for t in terms[1:]:
      s = s + t
    return np.where(np.isfinite(s), s, NEG_INF).astype(f32)


def hmc_one_step(target, x, lp, g, step_size, L, seed, layout=orng.PARTITIONABLE,
                 chain_offset=0, B_global=None, inv_mass=None):
  """HamiltonianMonteCarlo.one_step = MetropolisHastings(UncalibratedHMC).
  Returns dict with next state/results and the proposal details.  `inv_mass` [D] (diagonal inverse mass matrix =
  running variance): PreconditionedHamiltonianMonteCarlo with the momentum distribution of
  DiagonalMassMatrixAdaptation (experimental/mcmc/diagonal_mass_matrix_adaptation.py:73,
  MultivariateNormalPrecisionFactorLinearOperator with precision factor diag(sqrt(variance))): momentum = z / sqrt(var),
  kinetic energy 1/2 sum var m^2, velocity var m."""
  x = np.asarray(x, f32)
  B, D = x.shape
  prop, acc = orng.split(seed, 2, layout)                         # metropolis_hastings.py:183
  part_keys = orng.split(prop, len(target.part_sizes), layout)    # hmc.py:685
  m0 = draw_momentum(part_keys, B, target.part_sizes, layout, chain_offset, B_global)
  if inv_mass is not None:
    inv_mass = np.asarray(inv_mass, f32)
    m0 = (m0 / np.sqrt(inv_mass).astype(f32)).astype(f32)
  eps = _step_b(step_size, B, D)
  m1, x1, lp1, g1 = leapfrog(target, m0, x, lp, g, eps, L, inv_mass)
  with np.errstate(invalid='ignore', over='ignore'):
    w = f32(1) if inv_mass is None else inv_mass
    k0 = np.sum(w * m0 * m0, axis=1, dtype=f32)
    k1 = np.sum(w * m1 * m1, axis=1, dtype=f32)
    corr = f32(0.5) * safe_sum([k0, -k1])                         # hmc.py:862-875
    ratio = safe_sum([lp1, -np.asarray(lp, f32), corr])           # metropolis_hastings.py:204-215
    u = _uniform_b(acc, B, layout, chain_offset, B_global)
    with np.errstate(divide='ignore'):
      log_u = np.log(u).astype(f32)
    accept = log_u < ratio                                        # :221-227
  a = accept[:, None]
  return dict(
      state=np.where(a, x1, x), target_log_prob=np.where(accept, lp1, lp).astype(f32),
      grads=np.where(a, g1, g), is_accepted=accept, log_accept_ratio=ratio,
      proposed_state=x1, proposed_target_log_prob=lp1, proposed_grads=g1,
      log_acceptance_correction=corr, initial_momentum=m0, final_momentum=m1)


# ----------------------------------------------------------------------------
# NUTS instruction tables
def build_tree_uturn_instruction(max_depth, init_memory=0):
  """nuts.py:1013-1028: (left, right) leaf pairs of every balanced subtree."""
  out = set()

  def rec(addr, depth):
    if depth == 0:
      return addr + 1, addr + 1
    lft, rgt = rec(addr, depth - 1)
    _, rgt = rec(rgt, depth - 1)
    out.add((lft, rgt))
    return lft, rgt

  rec(init_memory, max_depth)
  return np.array(sorted(out), np.int32)


def write_read_instructions(max_tree_depth):
  """nuts.py:1031-1071 restated via its memory-footprint definition: a leaf is
  stored iff some later leaf checks against it (even steps); slot = number of
  live stored leaves; odd steps read the contiguous live range ending at the top."""
  instr = build_tree_uturn_instruction(max_tree_depth, init_memory=-1)
  n = int(instr.max()) + 1
  last_use = np.full(n, -1, np.int64)
  for a, b in instr:
    last_use[a] = max(last_use[a], b)
  nchecks = np.zeros(n, np.int64)
  for a, b in instr:
    nchecks[b] += 1
  write = np.zeros(n, np.int64)
  read = np.zeros((n, 2), np.int64)
  live_counts = np.zeros(n, np.int64)
  for i in range(n):
    live = [a for a in range(i + 1) if last_use[a] >= i]   # stored and still needed (incl. self)
    live_counts[i] = len(live)
  trash = int(live_counts.max())
  for i in range(n):
    if last_use[i] >= 0:
      write[i] = live_counts[i] - 1
    else:
      write[i] = trash
    if nchecks[i] > 0:
      live_prev = [a for a in range(i) if last_use[a] >= i]
      need = sorted(a for a, b in instr if b == i)
      first = live_prev.index(need[0])
      read[i] = (first, first + len(need))
  return write.astype(np.int32), read.astype(np.int32)


def write_read_closed_form(max_tree_depth):
  """Closed form used by the CUDA kernel: write[i] = popcount(i) (even i) or the
  trash slot max_tree_depth (odd i); read[i] = [popcount(i) - trailing_ones(i),
  popcount(i)) for odd i, [0,0) for even i."""
  n = 1 << max_tree_depth
  write = np.zeros(n, np.int32)
  read = np.zeros((n, 2), np.int32)
  for i in range(n):
    pc = bin(i).count('1')
    if i % 2 == 0:
      write[i] = pc
    else:
      write[i] = max_tree_depth
      t = 0
      while (i >> t) & 1:
        t += 1
      read[i] = (pc - t, pc)
  return write, read


# ----------------------------------------------------------------------------
def _dot(a, b):
  return np.sum(a * b, axis=1, dtype=f32)


def _energy(lp, m, inv_mass=None):
  """compute_hamiltonian nuts.py:1085-1102 (preconditioned: kinetic energy 1/2 sum inv_mass m^2, up to the constant of
  the momentum distribution's log-prob, which cancels in every energy difference)."""
  with np.errstate(invalid='ignore', over='ignore'):
    w = f32(1) if inv_mass is None else inv_mass
    return (lp - f32(0.5) * np.sum(w * m * m, axis=1, dtype=f32)).astype(f32)


def nuts_one_step(target, x, lp, g, step_size, seed, max_tree_depth=10,
                  max_energy_diff=1000.0, unrolled_leapfrog_steps=1,
                  layout=orng.PARTITIONABLE, chain_offset=0, B_global=None,
                  count_grad_evals=None, inv_mass=None):
  """Batched, lock-step NoUTurnSampler.one_step (nuts.py:321-445), literal masks.  `inv_mass` [D]: the
  PreconditionedNoUTurnSampler with a diagonal mass matrix (experimental/mcmc/preconditioned_nuts.py:169): momentum
  = z / sqrt(inv_mass), positions move along the velocity inv_mass * m, and the U-turn criterion dots the cumulative
  MOMENTUM with the end VELOCITIES (:694-705, :964-1030)."""
  with np.errstate(all='ignore'):   # stopped chains keep integrating in lock-step and may overflow
    return _nuts_one_step(target, x, lp, g, step_size, seed, max_tree_depth, max_energy_diff,
                          unrolled_leapfrog_steps, layout, chain_offset, B_global, count_grad_evals, inv_mass)


def _nuts_one_step(target, x, lp, g, step_size, seed, max_tree_depth, max_energy_diff,
                   unrolled_leapfrog_steps, layout, chain_offset, B_global, count_grad_evals, inv_mass=None):
  x = np.asarray(x, f32); lp = np.asarray(lp, f32); g = np.asarray(g, f32)
  B, D = x.shape
  write_instr, read_instr = write_read_closed_form(max_tree_depth)
  eps_abs = _step_b(step_size, B, D)
  k_start, k_loop = orng.split(seed, 2, layout)                         # :323
  ks = orng.split(k_start, len(target.part_sizes) + 1, layout)          # :515
  m = draw_momentum(ks[:-1], B, target.part_sizes, layout, chain_offset, B_global)
  vel = (lambda mm: mm) if inv_mass is None else (lambda mm: (np.asarray(inv_mass, f32) * mm).astype(f32))
  if inv_mass is not None:
    inv_mass = np.asarray(inv_mass, f32)
    m = (m / np.sqrt(inv_mass).astype(f32)).astype(f32)
  H0 = _energy(lp, m, inv_mass)                                          # :524
  # [2,B,...] ends: index 0 = left, 1 = right (:344-355)
  end_m = np.stack([m, m]); end_x = np.stack([x, x])
  end_lp = np.stack([lp, lp]); end_g = np.stack([g, g])
  cand = dict(x=x.copy(), lp=lp.copy(), g=g.copy(), energy=H0.copy(), w=np.zeros(B, f32))
  rho = m.copy()
  esum = np.zeros(B, f32)
  nleap = np.zeros(B, np.int32)
  cont = np.ones(B, bool)
  notdiv = np.ones(B, bool)
  accepted = np.zeros(B, bool)
  Sm = np.zeros((max_tree_depth + 1, B, D), f32)      # init_momentum_state_memory :447-454
  Srho = np.zeros((max_tree_depth + 1, B, D), f32)
  key = k_loop
  it = 0
  med = f32(max_energy_diff)
  while it < max_tree_depth and cont.any():                              # :404-407
    k_dir, k_sub, k_acc, key = orng.split(key, 4, layout)                # :546-549
    Bg = B if B_global is None else B_global
    direction = orng.randint_bit(k_dir, Bg, layout)[chain_offset:chain_offset + B].astype(bool)
    dcol = direction[:, None]
    s_m = np.where(dcol, end_m[1], end_m[0]); s_x = np.where(dcol, end_x[1], end_x[0])
    s_lp = np.where(direction, end_lp[1], end_lp[0]); s_g = np.where(dcol, end_g[1], end_g[0])
    eps = np.where(dcol, eps_abs, -eps_abs).astype(f32)                  # :571-574
    # --- _build_sub_tree :713-791
    sub = dict(x=s_x.copy(), lp=s_lp.copy(), g=s_g.copy(), energy=s_lp.copy(),
               w=np.full(B, NEG_INF, f32))
    rho_s = np.zeros((B, D), f32)
    n_sub = np.zeros(B, np.int32)
    c_prev = cont.copy()
    nd = notdiv.copy()
    kk = k_sub
    i = 0
    nsteps = 1 << it
    esum_sub = np.zeros(B, f32)
    while i < nsteps and c_prev.any():                                   # :759
      k_u, kk = orng.split(kk, 2, layout)                                # :808
      s_m, s_x, s_lp, s_g = leapfrog(target, s_m, s_x, s_lp, s_g, eps, unrolled_leapfrog_steps, inv_mass)
      if count_grad_evals is not None:
        count_grad_evals[0] += B * unrolled_leapfrog_steps
      rho_prev = rho_s
      rho_s = (rho_s + s_m).astype(f32)                                  # :826
      n_sub = np.where(c_prev, n_sub + 1, n_sub)                         # :829
      # U-turn checks against stored checkpoints (:846-855, :949-984)
      ok = np.ones(B, bool)
      r0, r1 = read_instr[i]
      for kidx in range(r0, r1):
        if not ok.any():
          break
        diff = (rho_s - Srho[kidx]).astype(f32)
        with np.errstate(invalid='ignore', over='ignore'):
          ok = ok & (_dot(diff, vel(Sm[kidx])) >= 0) & (_dot(diff, vel(s_m)) >= 0)
      w_i = write_instr[i]
      Sm[w_i] = s_m; Srho[w_i] = rho_prev                                # :859-869
      en = _energy(s_lp, s_m, inv_mass)
      en = np.where(np.isnan(en), NEG_INF, en).astype(f32)               # :874-876
      with np.errstate(invalid='ignore', over='ignore'):
        dH = (en - H0).astype(f32)
        nd_i = (-dH) < med                                               # :880
        w_new = log_add_exp(sub['w'], dH)
        thresh = (dH - w_new).astype(f32)
        u = _uniform_b(k_u, B, layout, chain_offset, B_global)
        take = np.log1p(-u).astype(f32) <= thresh                        # :897-901
      tcol = take[:, None]
      sub = dict(x=np.where(tcol, s_x, sub['x']), lp=np.where(take, s_lp, sub['lp']),
                 g=np.where(tcol, s_g, sub['g']), energy=np.where(take, en, sub['energy']),
                 w=w_new)
      c_now = nd_i & c_prev                                              # :921
      c_next = ok & c_now
      nd_keep = np.where(c_prev, nd_i, True)                             # :924-927
      with np.errstate(over='ignore', invalid='ignore'):
        ee = np.exp(np.minimum(dH, f32(0))).astype(f32)                  # :930
      esum_sub = np.where(c_now, esum_sub + ee, esum_sub).astype(f32)
      nd = nd & nd_keep
      c_prev = c_next
      i += 1
    cont_f = c_prev
    # --- back in _loop_tree_doubling :597-711
    esum = (esum_sub + esum).astype(f32)
    tw = np.where(cont_f, sub['w'], NEG_INF).astype(f32)
    with np.errstate(invalid='ignore'):
      wsum = log_add_exp(tw, cand['w'])
      thr = (tw - cand['w']).astype(f32)
    thr = np.where(np.isnan(thr), f32(0), thr)
    u = _uniform_b(k_acc, B, layout, chain_offset, B_global)
    swap = (np.log1p(-u).astype(f32) <= thr) & cont_f
    sc = swap[:, None]
    cand = dict(x=np.where(sc, sub['x'], cand['x']), lp=np.where(swap, sub['lp'], cand['lp']),
                g=np.where(sc, sub['g'], cand['g']),
                energy=np.where(swap, sub['energy'], cand['energy']), w=wsum)
    # ends: new [left, right] = [dir ? other : final, dir ? final : other]  (:664-675)
    o_m = np.where(dcol, end_m[0], end_m[1]); o_x = np.where(dcol, end_x[0], end_x[1])
    o_lp = np.where(direction, end_lp[0], end_lp[1]); o_g = np.where(dcol, end_g[0], end_g[1])
    end_m = np.stack([np.where(dcol, o_m, s_m), np.where(dcol, s_m, o_m)])
    end_x = np.stack([np.where(dcol, o_x, s_x), np.where(dcol, s_x, o_x)])
    end_lp = np.stack([np.where(direction, o_lp, s_lp), np.where(direction, s_lp, o_lp)])
    end_g = np.stack([np.where(dcol, o_g, s_g), np.where(dcol, s_g, o_g)])
    rho = (rho + rho_s).astype(f32)                                      # :677-682
    with np.errstate(invalid='ignore', over='ignore'):
      no_uturn = (_dot(rho, vel(end_m[0])) >= 0) & (_dot(rho, vel(end_m[1])) >= 0)  # :694-699
    accepted = swap | accepted
    cont = cont_f & no_uturn
    notdiv = nd
    nleap = nleap + n_sub
    it += 1
  with np.errstate(divide='ignore', invalid='ignore'):
    lar = np.log(esum / nleap.astype(f32)).astype(f32)                   # :429-432
  return dict(state=cand['x'], target_log_prob=cand['lp'], grads=cand['g'],
              energy=cand['energy'], log_accept_ratio=lar,
              leapfrogs_taken=(nleap * unrolled_leapfrog_steps).astype(np.int32),
              is_accepted=accepted, reach_max_depth=cont, has_divergence=~notdiv)


class _OneChain:
  """Adapter so per-chain code can call a batched target on a single row."""

  def __init__(self, target):
    self.t = target
    self.part_sizes = target.part_sizes

  def logp_grad(self, x):
    return self.t.logp_grad(x)


def nuts_one_step_chain(target, x, lp, g, step_size, seed, max_tree_depth=10,
                        max_energy_diff=1000.0, unrolled_leapfrog_steps=1,
                        layout=orng.PARTITIONABLE, chain_offset=0, B_global=None):
  """Per-chain early-exit NUTS: the form the warp-per-chain CUDA kernel
  follows.  Must agree EXACTLY with nuts_one_step (same keys, same counters)."""
  x = np.asarray(x, f32); lp = np.asarray(lp, f32); g = np.asarray(g, f32)
  B, D = x.shape
  Bg = B if B_global is None else B_global
  write_instr, read_instr = write_read_closed_form(max_tree_depth)
  eps_abs_all = _step_b(step_size, B, D)
  # ---- chain-independent key schedule
  k_start, k_loop = orng.split(seed, 2, layout)
  ks = orng.split(k_start, len(target.part_sizes) + 1, layout)
  m_all = draw_momentum(ks[:-1], B, target.part_sizes, layout, chain_offset, B_global)
  sched = []
  key = k_loop
  for it in range(max_tree_depth):
    k_dir, k_sub, k_acc, key = orng.split(key, 4, layout)
    dirs = orng.randint_bit(k_dir, Bg, layout)[chain_offset:chain_offset + B].astype(bool)
    uacc = _uniform_b(k_acc, B, layout, chain_offset, B_global)
    sched.append((dirs, k_sub, uacc))
  out = dict(state=np.empty_like(x), target_log_prob=np.empty(B, f32), grads=np.empty_like(x),
             energy=np.empty(B, f32), log_accept_ratio=np.empty(B, f32),
             leapfrogs_taken=np.empty(B, np.int32), is_accepted=np.empty(B, bool),
             reach_max_depth=np.empty(B, bool), has_divergence=np.empty(B, bool))
  med = f32(max_energy_diff)
  # subtree uniforms are drawn lazily per (depth, i) and cached for all chains
  ucache = {}

  def sub_uniform(it, i):
    if (it, i) not in ucache:
      kk = sched[it][1]
      # walk the split chain up to i (cache intermediate keys)
      j = 0
      while True:
        k_u, kk = orng.split(kk, 2, layout)
        if (it, j) not in ucache:
          ucache[(it, j)] = _uniform_b(k_u, B, layout, chain_offset, B_global)
        if j == i:
          break
        j += 1
    return ucache[(it, i)]

  tc = _OneChain(target)
  for c in range(B):
    eps_abs = eps_abs_all[c:c + 1]
    m = m_all[c:c + 1]
    xc, lpc, gc = x[c:c + 1], lp[c:c + 1], g[c:c + 1]
    H0 = _energy(lpc, m)
    L = [m, xc, lpc, gc]; R = [m, xc, lpc, gc]
    cand = [xc.copy(), lpc.copy(), gc.copy(), H0.copy(), f32(0.0)]
    rho = m.copy()
    esum = f32(0); nleap = 0; cont = True; notdiv = True; acc = False
    Sm = np.zeros((max_tree_depth + 1, 1, D), f32); Srho = np.zeros_like(Sm)
    for it in range(max_tree_depth):
      if not cont:
        break
      direction = bool(sched[it][0][c])
      s = list(R if direction else L)
      eps = (eps_abs if direction else -eps_abs).astype(f32)
      sub = [s[1].copy(), s[2].copy(), s[3].copy(), s[2].copy(), NEG_INF]
      rho_s = np.zeros((1, D), f32)
      n = 0; c_prev = cont; nd = notdiv
      esum_sub = f32(0)
      for i in range(1 << it):
        if not c_prev:
          break
        s = list(leapfrog(tc, s[0], s[1], s[2], s[3], eps, unrolled_leapfrog_steps))
        rho_prev = rho_s
        rho_s = (rho_s + s[0]).astype(f32)
        n += 1
        ok = True
        r0, r1 = read_instr[i]
        for kidx in range(r0, r1):
          if not ok:
            break
          diff = (rho_s - Srho[kidx]).astype(f32)
          with np.errstate(invalid='ignore', over='ignore'):
            ok = ok and bool(_dot(diff, Sm[kidx])[0] >= 0) and bool(_dot(diff, s[0])[0] >= 0)
        Sm[write_instr[i]] = s[0]; Srho[write_instr[i]] = rho_prev
        en = _energy(s[2], s[0])
        en = np.where(np.isnan(en), NEG_INF, en).astype(f32)
        with np.errstate(invalid='ignore', over='ignore'):
          dH = (en - H0).astype(f32)
          nd_i = bool(((-dH) < med)[0])
          w_new = log_add_exp(np.asarray([sub[4]], f32), dH)
          thresh = (dH - w_new).astype(f32)
          u = sub_uniform(it, i)[c]
          take = bool((np.log1p(-u).astype(f32) <= thresh)[0])
        if take:
          sub[0], sub[1], sub[2], sub[3] = s[1], s[2], s[3], en
        sub[4] = w_new[0]
        c_now = nd_i and c_prev
        if c_now:
          with np.errstate(over='ignore', invalid='ignore'):
            esum_sub = f32(esum_sub + np.exp(np.minimum(dH, f32(0))).astype(f32)[0])
        nd = nd and nd_i
        c_prev = ok and c_now
      cont_f = c_prev
      esum = f32(esum_sub + esum)
      tw = sub[4] if cont_f else NEG_INF
      with np.errstate(invalid='ignore'):
        wsum = log_add_exp(np.asarray([tw], f32), np.asarray([cand[4]], f32))[0]
        thr = f32(tw - cand[4])
      if np.isnan(thr):
        thr = f32(0)
      u = sched[it][2][c]
      swap = bool(np.log1p(-u).astype(f32) <= thr) and cont_f
      cand[4] = wsum
      if swap:
        cand[0], cand[1], cand[2], cand[3] = sub[0], sub[1], sub[2], sub[3]
      if direction:
        R = s
      else:
        L = s
      rho = (rho + rho_s).astype(f32)
      nleap += n
      acc = acc or swap
      notdiv = nd
      with np.errstate(invalid='ignore', over='ignore'):
        cont = cont_f and bool(_dot(rho, L[0])[0] >= 0) and bool(_dot(rho, R[0])[0] >= 0)
    out['state'][c] = cand[0][0]; out['target_log_prob'][c] = cand[1][0]
    out['grads'][c] = cand[2][0]; out['energy'][c] = cand[3][0]
    with np.errstate(divide='ignore', invalid='ignore'):
      out['log_accept_ratio'][c] = np.log(f32(esum) / f32(nleap))
    out['leapfrogs_taken'][c] = nleap * unrolled_leapfrog_steps
    out['is_accepted'][c] = acc; out['reach_max_depth'][c] = cont
    out['has_divergence'][c] = not notdiv
  return out


# ----------------------------------------------------------------------------
class DualAveraging:
  """dual_averaging_step_size_adaptation.py:419-475 (update), :543-609 (bootstrap),
  log_accept_prob getter hmc-like = min(0, finite-or(-inf)) simple_step_size_adaptation.py:42-48.
  Scalar (chain-shared) step size: reduce over ALL chains."""

  def __init__(self, step_size, num_adaptation_steps, target_accept_prob=0.75,
               exploration_shrinkage=0.05, step_count_smoothing=10.0, decay_rate=0.75,
               shrinkage_target=None):
    self.step_size = f32(step_size)
    self.n_adapt = int(num_adaptation_steps)
    self.target = f32(target_accept_prob)
    self.gamma = f32(exploration_shrinkage)
    self.t0 = f32(step_count_smoothing)
    self.kappa = f32(decay_rate)
    self.error_sum = f32(0)
    self.log_avg = f32(0)
    self.step = 0
    self.log_shrink = (f32(np.log(10.0)) + np.log(f32(step_size))).astype(f32) \
        if shrinkage_target is None else np.log(f32(shrinkage_target))

  @staticmethod
  def reduce_logmeanexp(la, extra_partials=None):
    """math/generic.py:221-274 via distribute_lib.reduce_logsumexp :147-162."""
    la = np.asarray(la, f32)
    xmax = np.max(la)
    xmax = xmax if np.isfinite(xmax) else f32(0)
    with np.errstate(divide='ignore'):
      lse = xmax + np.log(np.sum(np.exp(la - xmax), dtype=f32))
    return f32(lse - np.log(f32(la.size)))

  def update(self, log_accept_ratio):
    lar = np.asarray(log_accept_ratio, f32)
    la = np.minimum(np.where(np.isfinite(lar), lar, NEG_INF), f32(0)).astype(f32)
    r = self.reduce_logmeanexp(la)
    self.error_sum = f32(self.error_sum + self.target - np.exp(r))
    t = f32(self.step + 1)
    log_step = f32(self.log_shrink - (self.error_sum * np.sqrt(t)) / ((self.t0 + t) * self.gamma))
    eta = f32(t ** (-self.kappa))
    new_log_avg = f32(eta * log_step + (f32(1) - eta) * self.log_avg)
    step = self.step + 1
    if step < self.n_adapt:
      self.step_size = f32(np.exp(log_step))
    elif step == self.n_adapt:
      self.step_size = f32(np.exp(new_log_avg))
    if step <= self.n_adapt:
      self.log_avg = new_log_avg
    self.step = step
    return self.step_size


class SimpleAdaptation:
  """simple_step_size_adaptation.py:353-440 for one chain-shared scalar step size: the step is multiplied by
  (1 + adaptation_rate) when the log-mean-exp accept prob exceeds log(target), divided otherwise, for the first
  num_adaptation_steps steps (hmc_test.py:917-952 pins the count).  Same interface as DualAveraging."""

  def __init__(self, step_size, num_adaptation_steps, target_accept_prob=0.75, adaptation_rate=0.01):
    self.step_size = f32(step_size)
    self.n_adapt = int(num_adaptation_steps)
    self.log_target = np.log(f32(target_accept_prob))
    self.one_plus = f32(1) + f32(adaptation_rate)
    self.step = 0

  def update(self, log_accept_ratio):
    lar = np.asarray(log_accept_ratio, f32)
    la = np.minimum(np.where(np.isfinite(lar), lar, NEG_INF), f32(0)).astype(f32)
    r = DualAveraging.reduce_logmeanexp(la)
    if self.step < self.n_adapt:
      self.step_size = f32(self.step_size * self.one_plus) if r > self.log_target else \
          f32(self.step_size / self.one_plus)
    self.step += 1
    return self.step_size


# ----------------------------------------------------------------------------
def sample_chain(target, kind, x0, num_results, num_burnin_steps=0, num_steps_between_results=0,
                 step_size=0.1, num_leapfrog_steps=3, max_tree_depth=10, max_energy_diff=1000.0,
                 seed=17, dual_averaging=None, layout=orng.PARTITIONABLE, nuts_impl=nuts_one_step,
                 count_grad_evals=None):
  """sample.py:311-383: salt seed, bootstrap, then per result (1+burnin | 1+thin) steps,
  each `step_seed, seed = split(seed)`."""
  seed = orng.sanitize_seed(seed, salt='mcmc.sample_chain')
  x = np.asarray(x0, f32)
  lp, g = target.logp_grad(x)
  states, trace = [], []
  eps = f32(step_size) if dual_averaging is None else dual_averaging.step_size
  for r in range(num_results):
    nsteps = 1 + (num_burnin_steps if r == 0 else num_steps_between_results)
    for _ in range(nsteps):
      step_seed, seed = orng.split(seed, 2, layout)
      if kind == 'hmc':
        res = hmc_one_step(target, x, lp, g, eps, num_leapfrog_steps, step_seed, layout)
        if count_grad_evals is not None:
          count_grad_evals[0] += x.shape[0] * num_leapfrog_steps
      else:
        kw = dict(count_grad_evals=count_grad_evals) if nuts_impl is nuts_one_step else {}
        res = nuts_impl(target, x, lp, g, eps, step_seed, max_tree_depth, max_energy_diff,
                        layout=layout, **kw)
      res['step_size'] = eps
      x, lp, g = res['state'], res['target_log_prob'], res['grads']
      if dual_averaging is not None:
        eps = dual_averaging.update(res['log_accept_ratio'])
    states.append(x.copy())
    trace.append({k: v for k, v in res.items() if k not in ('grads', 'proposed_grads')})
  return np.stack(states), trace, seed
