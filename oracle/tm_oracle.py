"""CPU oracle for the timemachine force-evaluation + integration hot path.

*** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under timemachine_b200/ does.

It is a NumPy float64 restatement of the reference's own CPU (JAX) potentials and integrator, with analytic gradients
(jax.grad is unavailable here: jax is not installed in this image), each function citing the reference lines it
follows (paths relative to /root/reference):

  nonbonded energy     timemachine/potentials/nonbonded.py:221-339 (`nonbonded`), switch_fn :23-39,
                       pair form :342-400, block form :82-150; PBC timemachine/potentials/jax_utils.py:37-44,144-181
  gradients / du_dp    analytic forms of timemachine/cpp/src/kernels/k_nonbonded_common.cuh:72-94,184-246 and
                       k_nonbonded.cuh:239-264 (which parameter each term feeds)
  bonded               timemachine/potentials/bonded.py:34-79 (bond), :82-138 (angle), :141-216 (torsion); gradients
                       as in kernels/k_harmonic_bond.cuh:40-53, k_harmonic_angle.cuh:76-141, k_periodic_torsion.cuh:72-126
  integrator           timemachine/integrator.py:15-53 (langevin_coefficients), :137-144 (BAOAB step); the
                       mixed-precision variant restates kernels/k_integrator.cuh:32-46
  neighbour list       tests/test_nblist.py:28-55 (block bounds), :117-139 (brute-force tile membership)
  hilbert              timemachine/cpp/src/kernels/k_hilbert.cu:19-47 (binning), hilbert_sort.cu:17-33 (LUT);
                       the curve itself is checked against the vendored C routine (oracle/_ref/libhilbert_ref.so)

PINNING: the energies of this oracle are checked against outputs of the reference's own Python functions, imported
from /root/reference with a numpy stand-in for jax.numpy (tests/golden/make_golden.py wrote tests/golden/*.npz; the
reference has no golden vectors of its own for this path), and its analytic gradients are checked
against central finite differences of those energies (tests/test_oracle.py).
"""

from __future__ import annotations

import numpy as np
from scipy.special import erfc

FIXED_EXPONENT = 1 << 36
BOLTZ = 0.008314462618  # timemachine/cpp/src/constants.hpp:5 (what the CUDA integrator uses)
SWITCH_CUTOFF = 1.2


# ---------------------------------------------------------------------------------------------------------------------
# geometry
def delta_r(ri, rj, box=None):
    """jax_utils.py:37-44"""
    diff = ri - rj
    if box is not None:
        bd = np.diag(box)
        diff = diff - bd * np.floor(diff / bd + 0.5)
    return diff


def switch_fn(d):
    """nonbonded.py:23-39; the 1.2 nm is intentionally hard-coded in the reference"""
    f = np.cos(0.5 * np.pi * (d / SWITCH_CUTOFF) ** 8) ** 3
    return np.where(d < SWITCH_CUTOFF, f, 0.0)


def d_switch_fn(d):
    arg = 0.5 * np.pi * (d / SWITCH_CUTOFF) ** 8
    g = -12.0 * np.pi * d**7 / SWITCH_CUTOFF**8 * np.sin(arg) * np.cos(arg) ** 2
    return np.where(d < SWITCH_CUTOFF, g, 0.0)


def _pair_terms(d, qi, qj, si, sj, ei, ej, beta, q_scale=1.0, lj_scale=1.0):
    """Energy and derivatives of one pair at 4-D distance d (arrays broadcast).

    Returns dict with u, du_dd, and the du/dparam pieces:  dq_i, dq_j, dsig (same for i and j), deps_i, deps_j.
    """
    inv_d = 1.0 / d
    bd = beta * d
    e = erfc(bd)
    de = -2.0 * beta / np.sqrt(np.pi) * np.exp(-bd * bd)
    s = switch_fn(d)
    ds = d_switch_fn(d)
    damp = e * s
    ddamp = e * ds + de * s
    qij = q_scale * qi * qj
    u_es = qij * damp * inv_d
    du_es = qij * (ddamp * inv_d - damp * inv_d * inv_d)

    lj_on = (ei != 0) & (ej != 0)
    eps_ij = np.where(lj_on, ei * ej, 0.0)
    sig_ij = si + sj
    s6 = (sig_ij * inv_d) ** 6
    s12 = s6 * s6
    u_lj = lj_scale * 4.0 * eps_ij * (s12 - s6)
    du_lj = -lj_scale * 24.0 * eps_ij * inv_d * (2.0 * s12 - s6)
    with np.errstate(divide="ignore", invalid="ignore"):
        dsig = np.where(lj_on, lj_scale * 24.0 * eps_ij * (2.0 * s12 - s6) / sig_ij, 0.0)
    deps_common = np.where(lj_on, lj_scale * 4.0 * (s12 - s6), 0.0)
    return {
        "u": u_es + u_lj,
        "du_dd": du_es + du_lj,
        "dq_i": q_scale * qj * damp * inv_d,
        "dq_j": q_scale * qi * damp * inv_d,
        "dsig": dsig,
        "deps_i": deps_common * ej,
        "deps_j": deps_common * ei,
    }


# ---------------------------------------------------------------------------------------------------------------------
# nonbonded: rows x cols block with optional pair mask, the building block of every variant
def _nonbonded_block(x, params, box, rows, cols, beta, cutoff, upper_triangular, q_mask=None, lj_mask=None, chunk=256):
    """Accumulate u, du_dx[N,3], du_dp[N,4] over pairs (i in rows, j in cols); if upper_triangular only i < j (by atom
    index) counts. q_mask / lj_mask: optional dense [N,N] multipliers (exclusion rescale masks, nonbonded.py:159-173)."""
    N = x.shape[0]
    du_dx = np.zeros((N, 3))
    du_dp = np.zeros((N, 4))
    u_total = 0.0
    rows = np.asarray(rows)
    cols = np.asarray(cols)
    xj = x[cols]
    pj = params[cols]
    for start in range(0, len(rows), chunk):
        r = rows[start : start + chunk]
        xi = x[r]
        pi = params[r]
        dxyz = delta_r(xi[:, None, :], xj[None, :, :], box)  # (R, C, 3)
        dw = pi[:, None, 3] - pj[None, :, 3]
        d2 = np.sum(dxyz * dxyz, axis=-1) + dw * dw
        keep = d2 < cutoff * cutoff
        if upper_triangular:
            keep &= r[:, None] < cols[None, :]
        else:
            keep &= r[:, None] != cols[None, :]
        d = np.sqrt(np.where(keep, d2, 1.0))
        qs = 1.0 if q_mask is None else q_mask[np.ix_(r, cols)]
        ls = 1.0 if lj_mask is None else lj_mask[np.ix_(r, cols)]
        t = _pair_terms(d, pi[:, None, 0], pj[None, :, 0], pi[:, None, 1], pj[None, :, 1], pi[:, None, 2], pj[None, :, 2], beta, qs, ls)
        k = keep.astype(np.float64)
        u_total += float(np.sum(t["u"] * k))
        pref = t["du_dd"] / d * k  # multiplies the displacement
        f = pref[:, :, None] * dxyz
        np.add.at(du_dx, r, f.sum(axis=1))
        np.add.at(du_dx, cols, -f.sum(axis=0))
        gw = pref * dw
        np.add.at(du_dp[:, 3], r, gw.sum(axis=1))
        np.add.at(du_dp[:, 3], cols, -gw.sum(axis=0))
        np.add.at(du_dp[:, 0], r, (t["dq_i"] * k).sum(axis=1))
        np.add.at(du_dp[:, 0], cols, (t["dq_j"] * k).sum(axis=0))
        np.add.at(du_dp[:, 1], r, (t["dsig"] * k).sum(axis=1))
        np.add.at(du_dp[:, 1], cols, (t["dsig"] * k).sum(axis=0))
        np.add.at(du_dp[:, 2], r, (t["deps_i"] * k).sum(axis=1))
        np.add.at(du_dp[:, 2], cols, (t["deps_j"] * k).sum(axis=0))
    return u_total, du_dx, du_dp


def nonbonded_all_pairs(x, params, box, beta, cutoff, atom_idxs=None):
    """NonbondedAllPairs (no exclusions): every pair i<j of atom_idxs (potentials.py:126-138 decomposition)."""
    N = x.shape[0]
    idxs = np.arange(N) if atom_idxs is None else np.asarray(atom_idxs)
    return _nonbonded_block(x, params, box, idxs, idxs, beta, cutoff, upper_triangular=True)


def nonbonded_interaction_group(x, params, box, row_idxs, col_idxs, beta, cutoff):
    """rows x cols block, disjoint sets (nonbonded.py:82-150 nonbonded_block)."""
    return _nonbonded_block(x, params, box, np.asarray(row_idxs), np.asarray(col_idxs), beta, cutoff, upper_triangular=False)


def nonbonded_pair_list(x, params, box, pair_idxs, scales, beta, cutoff):
    """Explicit pairs with (charge, lj) rescale (nonbonded.py:342-400 nonbonded_on_specific_pairs)."""
    N = x.shape[0]
    du_dx = np.zeros((N, 3))
    du_dp = np.zeros((N, 4))
    pair_idxs = np.asarray(pair_idxs).reshape(-1, 2)
    scales = np.asarray(scales, dtype=np.float64).reshape(-1, 2)
    if len(pair_idxs) == 0:
        return 0.0, du_dx, du_dp
    i, j = pair_idxs[:, 0], pair_idxs[:, 1]
    dxyz = delta_r(x[i], x[j], box)
    dw = params[i, 3] - params[j, 3]
    d2 = np.sum(dxyz * dxyz, axis=-1) + dw * dw
    keep = d2 < cutoff * cutoff
    d = np.sqrt(np.where(keep, d2, 1.0))
    t = _pair_terms(d, params[i, 0], params[j, 0], params[i, 1], params[j, 1], params[i, 2], params[j, 2], beta, scales[:, 0], scales[:, 1])
    k = keep.astype(np.float64)
    u = float(np.sum(t["u"] * k))
    pref = t["du_dd"] / d * k
    f = pref[:, None] * dxyz
    np.add.at(du_dx, i, f)
    np.add.at(du_dx, j, -f)
    np.add.at(du_dp[:, 3], i, pref * dw)
    np.add.at(du_dp[:, 3], j, -pref * dw)
    np.add.at(du_dp[:, 0], i, t["dq_i"] * k)
    np.add.at(du_dp[:, 0], j, t["dq_j"] * k)
    np.add.at(du_dp[:, 1], i, t["dsig"] * k)
    np.add.at(du_dp[:, 1], j, t["dsig"] * k)
    np.add.at(du_dp[:, 2], i, t["deps_i"] * k)
    np.add.at(du_dp[:, 2], j, t["deps_j"] * k)
    return u, du_dx, du_dp


def nonbonded(x, params, box, exclusion_idxs, scale_factors, beta, cutoff, atom_idxs=None):
    """The monolithic `Nonbonded` potential: all pairs minus scaled exclusions (nonbonded.py:221-339;
    potentials.py:126-138 shows the GPU decomposition AllPairs + Exclusions that this equals)."""
    u, dx, dp = nonbonded_all_pairs(x, params, box, beta, cutoff, atom_idxs)
    exclusion_idxs = np.asarray(exclusion_idxs).reshape(-1, 2)
    scale_factors = np.asarray(scale_factors, dtype=np.float64).reshape(-1, 2)
    if atom_idxs is not None and len(exclusion_idxs):
        s = set(int(a) for a in atom_idxs)
        keep = np.array([int(i) in s and int(j) in s for i, j in exclusion_idxs], dtype=bool)
        exclusion_idxs, scale_factors = exclusion_idxs[keep], scale_factors[keep]
    ue, dxe, dpe = nonbonded_pair_list(x, params, box, exclusion_idxs, scale_factors, beta, cutoff)
    return u - ue, dx - dxe, dp - dpe


def nonbonded_energy_dense(x, params, box, exclusion_idxs, scale_factors, beta, cutoff):
    """Literal dense restatement of nonbonded.py:221-339 (energy only) - used to cross-check the blocked version and
    as the closest analogue of the reference's JAX CPU path for timing."""
    N = x.shape[0]
    q_mask = np.ones((N, N))
    lj_mask = np.ones((N, N))
    for (i, j), (qs, ls) in zip(np.asarray(exclusion_idxs).reshape(-1, 2), np.asarray(scale_factors).reshape(-1, 2)):
        q_mask[i, j] = q_mask[j, i] = 1 - qs
        lj_mask[i, j] = lj_mask[j, i] = 1 - ls
    q, sig, eps, w = params.T
    d_ijk = delta_r(x[:, None], x[None, :], box)
    d2 = np.sum(d_ijk**2, axis=2) + (w[:, None] - w[None, :]) ** 2
    np.fill_diagonal(d2, 0.0)
    dij = np.sqrt(d2)
    eye = np.eye(N, dtype=bool)
    eps_ij = eps[:, None] * eps[None, :]
    sig_ij = sig[:, None] + sig[None, :]
    keep = (~eye) & (eps_ij != 0) & (dij < cutoff)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(eye, 0.0, 1.0 / np.where(eye, 1.0, dij))
        s6 = (np.where(keep, sig_ij, 0.0) * inv) ** 6
        e_lj = np.where(keep, 4 * eps_ij * (s6 - 1.0) * s6, 0.0)
        qij = np.where(eye, 0.0, q[:, None] * q[None, :])
        e_q = np.where(eye | (dij >= cutoff), 0.0, qij * erfc(beta * dij) * inv * switch_fn(dij))
    return float(np.sum((e_lj * lj_mask + e_q * q_mask) / 2))


# ---------------------------------------------------------------------------------------------------------------------
# bonded
def harmonic_bond(x, params, bond_idxs):
    """bonded.py:34-79; returns (u, du_dx, du_dp[B,2])"""
    N = x.shape[0]
    du_dx = np.zeros((N, 3))
    bond_idxs = np.asarray(bond_idxs).reshape(-1, 2)
    params = np.asarray(params, dtype=np.float64).reshape(-1, 2)
    if len(bond_idxs) == 0:
        return 0.0, du_dx, np.zeros_like(params)
    i, j = bond_idxs[:, 0], bond_idxs[:, 1]
    dx = x[i] - x[j]
    r = np.sqrt(np.sum(dx * dx, axis=-1))
    kb, b0 = params[:, 0], params[:, 1]
    db = r - b0
    u = float(np.sum(kb / 2 * db * db))
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.where((b0 != 0)[:, None], (kb * db / r)[:, None] * dx, kb[:, None] * dx)
    np.add.at(du_dx, i, g)
    np.add.at(du_dx, j, -g)
    du_dp = np.stack([0.5 * db * db, -kb * db], axis=1)
    return u, du_dx, du_dp


def _angle_theta(x, angle_idxs, eps):
    i, j, k = angle_idxs[:, 0], angle_idxs[:, 1], angle_idxs[:, 2]
    rji = np.concatenate([x[i] - x[j], eps[:, None]], axis=1)
    rjk = np.concatenate([x[k] - x[j], eps[:, None]], axis=1)
    nji = np.linalg.norm(rji, axis=1, keepdims=True)
    njk = np.linalg.norm(rjk, axis=1, keepdims=True)
    y = np.linalg.norm(njk * rji - nji * rjk, axis=1)
    xx = np.linalg.norm(njk * rji + nji * rjk, axis=1)
    return 2 * np.arctan2(y, xx), rji, rjk, nji[:, 0], njk[:, 0]


def harmonic_angle(x, params, angle_idxs):
    """bonded.py:82-138 (Kahan-stable angle with the eps 4th component); returns (u, du_dx, du_dp[A,3])"""
    N = x.shape[0]
    du_dx = np.zeros((N, 3))
    angle_idxs = np.asarray(angle_idxs).reshape(-1, 3)
    params = np.asarray(params, dtype=np.float64).reshape(-1, 3)
    if len(angle_idxs) == 0:
        return 0.0, du_dx, np.zeros_like(params)
    ka, a0, eps = params[:, 0], params[:, 1], params[:, 2]
    theta, a, b, na, nb = _angle_theta(x, angle_idxs, eps)
    delta = theta - a0
    u = float(np.sum(ka / 2 * delta * delta))
    ab = np.sum(a * b, axis=1, keepdims=True)
    aa = np.sum(a * a, axis=1, keepdims=True)
    bb = np.sum(b * b, axis=1, keepdims=True)
    aab = a * ab - b * aa
    bba = b * ab - a * bb
    aab_n = np.linalg.norm(aab, axis=1, keepdims=True)
    bba_n = np.linalg.norm(bba, axis=1, keepdims=True)
    pref = (ka * delta)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        gi4 = np.where(aab_n == 0, 0.0, pref / na[:, None] * aab / aab_n)
        gk4 = np.where(bba_n == 0, 0.0, pref / nb[:, None] * bba / bba_n)
    i, j, k = angle_idxs[:, 0], angle_idxs[:, 1], angle_idxs[:, 2]
    np.add.at(du_dx, i, gi4[:, :3])
    np.add.at(du_dx, k, gk4[:, :3])
    np.add.at(du_dx, j, -gi4[:, :3] - gk4[:, :3])
    du_dp = np.stack([delta * delta / 2, -delta * ka, gi4[:, 3] + gk4[:, 3]], axis=1)
    return u, du_dx, du_dp


def periodic_torsion(x, params, torsion_idxs):
    """bonded.py:141-216; returns (u, du_dx, du_dp[T,3])"""
    N = x.shape[0]
    du_dx = np.zeros((N, 3))
    torsion_idxs = np.asarray(torsion_idxs).reshape(-1, 4)
    params = np.asarray(params, dtype=np.float64).reshape(-1, 3)
    if len(torsion_idxs) == 0:
        return 0.0, du_dx, np.zeros_like(params)
    i, j, k, l = torsion_idxs.T
    rij = x[j] - x[i]
    rkj = x[j] - x[k]
    rkl = x[l] - x[k]
    n1 = np.cross(rij, rkj)
    n2 = np.cross(rkj, rkl)
    rkj_n = np.linalg.norm(rkj, axis=1, keepdims=True)
    y = np.sum(np.cross(n1, n2) * rkj / rkj_n, axis=1)
    xx = np.sum(n1 * n2, axis=1)
    phi = np.arctan2(y, xx)
    kt, phase, period = params[:, 0], params[:, 1], params[:, 2]
    arg = period * phi - phase
    u = float(np.sum(kt * (1 + np.cos(arg))))
    n1_2 = np.sum(n1 * n1, axis=1, keepdims=True)
    n2_2 = np.sum(n2 * n2, axis=1, keepdims=True)
    rkj2 = rkj_n * rkj_n
    d0 = rkj_n / n1_2 * n1
    d3 = -rkj_n / n2_2 * n2
    ij_kj = np.sum(rij * rkj, axis=1, keepdims=True)
    kl_kj = np.sum(rkl * rkj, axis=1, keepdims=True)
    d1 = (ij_kj / rkj2 - 1) * d0 - d3 * kl_kj / rkj2
    d2 = (kl_kj / rkj2 - 1) * d3 - d0 * ij_kj / rkj2
    pref = (kt * np.sin(arg) * period)[:, None]  # note: du/dphi = -k n sin(n phi - phi0); d* hold -dphi/dx
    np.add.at(du_dx, i, d0 * pref)
    np.add.at(du_dx, j, d1 * pref)
    np.add.at(du_dx, k, d2 * pref)
    np.add.at(du_dx, l, d3 * pref)
    du_dp = np.stack([1 + np.cos(arg), kt * np.sin(arg), -kt * np.sin(arg) * phi], axis=1)
    return u, du_dx, du_dp


# ---------------------------------------------------------------------------------------------------------------------
# integrator
def langevin_coefficients(temperature, dt, friction, masses, boltz=BOLTZ):
    """integrator.py:15-53"""
    masses = np.asarray(masses, dtype=np.float64)
    kT = boltz * temperature
    ca = np.exp(-friction * dt)
    cb = dt / masses
    cc = np.sqrt(1 - np.exp(-2 * friction * dt)) * np.sqrt(kT / masses)
    return ca, cb, cc


def baoab_step(x, v, force, ca, cb, cc, dt, noise):
    """integrator.py:137-144 (_step), float64 throughout"""
    v_mid = v + cb[:, None] * force
    new_v = ca * v_mid + cc[:, None] * noise
    new_x = x + 0.5 * dt * (v_mid + new_v)
    return new_x, new_v


def velocity_verlet_multiple_steps(force_fn, x, v, masses, dt, n_steps):
    """integrator.py:169-201 (VelocityVerletIntegrator.multiple_steps): x and v are carried in 64-bit fixed point, every
    increment is rounded to fixed point before it is added.  force_fn(x) = -du/dx.  Returns (xs, vs), n_steps + 1 entries:
    the start, the state after each of the n_steps - 1 inner steps (velocities half a step behind), the end."""
    cb = dt / np.asarray(masses, dtype=np.float64)[:, None]
    # jnp.int64(v * 2^36) truncates toward zero (lib/fixed_point.py:14-16)
    to_fixed = lambda a: np.trunc(np.asarray(a, dtype=np.float64) * FIXED_EXPONENT).astype(np.int64)  # noqa: E731
    to_float = lambda a: a.astype(np.float64) / FIXED_EXPONENT  # noqa: E731
    x_fixed, v_fixed = to_fixed(x), to_fixed(v)
    zs = [(x_fixed, v_fixed)]
    v_fixed = v_fixed + to_fixed((0.5 * cb) * force_fn(to_float(x_fixed)))
    x_fixed = x_fixed + to_fixed(dt * to_float(v_fixed))
    for _ in range(n_steps - 1):
        v_fixed = v_fixed + to_fixed(cb * force_fn(to_float(x_fixed)))
        x_fixed = x_fixed + to_fixed(dt * to_float(v_fixed))
        zs.append((x_fixed, v_fixed))
    v_fixed = v_fixed + to_fixed((0.5 * cb) * force_fn(to_float(x_fixed)))
    zs.append((x_fixed, v_fixed))
    return to_float(np.array([a for a, _ in zs])), to_float(np.array([b for _, b in zs]))


def velocity_verlet_f64(du_dx_fn, x, v, cbs, dt, n_steps):
    """The compiled path (verlet_integrator.cu:26-110, k_integrator.cuh:64-130): plain f64 state, cbs = -dt / m,
    initialize = half kick + drift, n_steps x (kick + drift), finalize = half kick.  du_dx_fn(x) = +du/dx as the fixed
    point buffer holds it.  Returns (frames after every step_fwd, final x, final v)."""
    x = np.array(x, dtype=np.float64)
    v = np.array(v, dtype=np.float64)
    cbs = np.asarray(cbs, dtype=np.float64)[:, None]
    v = v + (0.5 * cbs) * du_dx_fn(x)
    x = x + dt * v
    frames = []
    for _ in range(n_steps):
        v = v + cbs * du_dx_fn(x)
        x = x + dt * v
        frames.append(x.copy())
    v = v + (0.5 * cbs) * du_dx_fn(x)
    return np.array(frames), x, v


def float_to_fixed(v):
    """k_fixed_point.cuh:10-24 is round-half-even of v * 2^36 (SURVEY.md §8c sub-oracle 3)"""
    return np.rint(np.asarray(v, dtype=np.float64) * FIXED_EXPONENT).astype(np.int64).view(np.uint64)


def fixed_to_float(v):
    return np.asarray(v, dtype=np.uint64).view(np.int64).astype(np.float64) / FIXED_EXPONENT


def fma32_exact(a, b, c):
    """Correctly rounded float32 fma(a, b, c), vectorised.  The product of two floats is exact in double and so is the
    error of the double sum (two-sum); rounding that sum to float can only go wrong when it sits exactly on a float
    midpoint while the discarded error is non-zero - then the error's sign decides instead of ties-to-even."""
    a64, b64, c64 = (np.asarray(t, dtype=np.float32).astype(np.float64) for t in (a, b, c))
    p = a64 * b64
    s = p + c64
    bb = s - p
    err = (p - (s - bb)) + (c64 - bb)
    r = s.astype(np.float32)
    r64 = r.astype(np.float64)
    up = np.nextafter(r, np.float32(np.inf)).astype(np.float64)
    dn = np.nextafter(r, np.float32(-np.inf)).astype(np.float64)
    d = s - r64
    with np.errstate(invalid="ignore"):
        tie_above = (d > 0) & (d == 0.5 * (up - r64))  # s is the midpoint of [r, up]: ties-to-even chose r
        tie_below = (d < 0) & (d == -0.5 * (r64 - dn))  # s is the midpoint of [dn, r]
    out = r.copy()
    out = np.where(tie_above & (err > 0), up.astype(np.float32), out)
    out = np.where(tie_below & (err < 0), dn.astype(np.float32), out)
    return out.astype(np.float32)


def baoab_step_mixed(x, v, du_dx_fixed, masses, temperature, dt, friction, noise_f32):
    """Bit-level restatement of k_integrator.cuh:32-46 + langevin_integrator.cu:17-30 AS COMPILED (nvcc contracts
    `ca * v_mid + ccs * noise` into fma(v_mid, ca, ccs * noise): FFMA in the SASS of the reference's
    k_update_forward_baoab<float>): f32 coefficients/force, f64 state.  du_dx_fixed: uint64[N,3]."""
    dt32 = np.float32(dt)
    ca = np.float32(np.exp(-friction * dt))
    kT = BOLTZ * temperature
    adj = np.sqrt(1 - np.exp(-2 * friction * dt))
    masses = np.asarray(masses, dtype=np.float64)
    cbs = (np.float64(dt32) / masses).astype(np.float32)
    ccs = (adj * np.sqrt(kT / masses)).astype(np.float32)
    f = -(np.asarray(du_dx_fixed, dtype=np.uint64).view(np.int64).astype(np.float32) / np.float32(FIXED_EXPONENT))
    cbf = (cbs[:, None] * f).astype(np.float32)
    v_mid = (v + cbf.astype(np.float64)).astype(np.float32)
    v_new = fma32_exact(np.full_like(v_mid, ca), v_mid, (ccs[:, None] * np.asarray(noise_f32, dtype=np.float32)).astype(np.float32))
    half_dt = np.float32(np.float32(0.5) * dt32)
    x_new = x + np.float64(half_dt) * (v_mid.astype(np.float64) + v_new.astype(np.float64))
    return x_new, v_new.astype(np.float64)


# ---------------------------------------------------------------------------------------------------------------------
# neighbour list
def reference_block_bounds(coords, box, block_size=32):
    """tests/test_nblist.py:28-55"""
    coords = np.array(coords, dtype=np.float64, copy=True)
    N = coords.shape[0]
    nb = (N + block_size - 1) // block_size
    bd = np.diagonal(box)
    ctrs, exts = [], []
    for b in range(nb):
        blk = coords[b * block_size : min((b + 1) * block_size, N)]
        lo = blk[0].copy()
        hi = blk[0].copy()
        for c in blk[1:]:
            center = 0.5 * (hi + lo)
            c = c - bd * np.floor((c - center) / bd + 0.5)
            lo = np.minimum(lo, c)
            hi = np.maximum(hi, c)
        ctrs.append((hi + lo) / 2)
        exts.append((hi - lo) / 2)
    return np.array(ctrs), np.array(exts)


def reference_ixn_list(coords, box, cutoff, block_size=32, row_idxs=None):
    """Brute-force canonical tile membership (tests/test_nblist.py:117-139, row-subset form :142-177).
    Returns (list of sorted j-lists per row block, min |d - cutoff| over all tested pairs)."""
    coords = np.asarray(coords, dtype=np.float64)
    N = coords.shape[0]
    out = []
    margin = np.inf
    if row_idxs is None:
        rows = np.arange(N)
        cols = np.arange(N)
        tri = True
    else:
        rows = np.asarray(row_idxs)
        cols = np.setdiff1d(np.arange(N), rows)
        tri = False
    nb = (len(rows) + block_size - 1) // block_size
    for b in range(nb):
        r = rows[b * block_size : (b + 1) * block_size]
        d = np.linalg.norm(delta_r(coords[r][:, None, :], coords[cols][None, :, :], box), axis=-1)
        if tri:
            d[:, : b * block_size] = np.inf
        margin = min(margin, float(np.min(np.abs(d - cutoff))))
        hit = np.any(d < cutoff, axis=0)
        out.append(sorted(cols[hit].tolist()))
    return out, margin


# ---------------------------------------------------------------------------------------------------------------------
# hilbert
def hilbert3d_index(c0, c1, c2, nbits=8):
    """Vectorised Butz/Moore index; exhaustively compared with the vendored C routine in tests/test_hilbert.py."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    index = np.zeros_like(c0)
    rot = np.zeros_like(c0)
    flip = np.zeros_like(c0)
    above = np.zeros_like(c0)
    one = np.uint64(1)
    for level in range(nbits - 1, -1, -1):
        lv = np.uint64(level)
        raw = (((c2 >> lv) & one) << np.uint64(2)) | (((c1 >> lv) & one) << one) | ((c0 >> lv) & one)
        digit = (raw ^ above) ^ flip
        digit = ((digit >> rot) | (digit << (np.uint64(3) - rot))) & np.uint64(7)
        index = (index << np.uint64(3)) | digit
        above = raw
        flip = one << rot
        low = digit & (~digit + one) & np.uint64(3)
        rot = rot + one + np.where(low == 1, 1, np.where(low == 2, 2, 0)).astype(np.uint64)
        rot = rot % np.uint64(3)
    total = 3 * nbits
    every_third = 0
    for b in range(0, total, 3):
        every_third |= 1 << b
    index ^= np.uint64(every_third >> 1)
    d = 1
    while d < total:
        index ^= index >> np.uint64(d)
        d *= 2
    return index


def hilbert_keys(coords, box, atom_idxs=None, grid=128):
    """k_hilbert.cu:19-47: home-box image with floor, bin = (unsigned)(x * min(1/b) * 127), key = curve index"""
    coords = np.asarray(coords, dtype=np.float64)
    if atom_idxs is not None:
        coords = coords[np.asarray(atom_idxs)]
    bd = np.diagonal(box).astype(np.float64)
    inv = 1.0 / bd
    inv_bin_width = np.min(inv) * (grid - 1.0)
    x = coords - bd * np.floor(coords * inv)
    bins = (x * inv_bin_width).astype(np.uint32)
    return hilbert3d_index(bins[:, 0], bins[:, 1], bins[:, 2]).astype(np.uint32)


def hilbert_perm(coords, box, atom_idxs=None):
    """hilbert_sort.cu:51-81: stable sort of (key, atom) pairs"""
    keys = hilbert_keys(coords, box, atom_idxs)
    order = np.argsort(keys, kind="stable")
    idxs = np.arange(len(coords)) if atom_idxs is None else np.asarray(atom_idxs)
    return idxs[order].astype(np.uint32)


# ---------------------------------------------------------------------------------------------------------------------
# Monte Carlo barostat.  Restates timemachine/cpp/src/kernels/k_barostat.cuh:11-189 and the
# host logic of timemachine/cpp/src/barostat.cu:153-250; the centroid scaling itself is pinned against the reference's
# Python CentroidRescaler (timemachine/md/barostat/moves.py:40-87) through tests/golden/barostat.npz.
AVOGADRO = 6.0221367e23  # timemachine/cpp/src/constants.hpp:6


def barostat_scale_centroids(coords, group_idxs, center, scale):
    """Rigidly move every group so that its centroid is scaled about `center` (float64; no re-imaging).
    Reference: CentroidRescaler.scale_centroids, md/barostat/moves.py:72-87."""
    out = np.array(coords, dtype=np.float64, copy=True)
    for g in group_idxs:
        g = np.asarray(g)
        centroid = out[g].mean(axis=0)
        out[g] += (center + scale * (centroid - center)) - centroid
    return out


def _f32(v):
    return np.float32(v)


def _fma32(a, b, c):
    # float32 fused multiply-add: the double product of two floats is exact; the final double->float rounding can
    # differ from a true fma only in double-rounding corner cases (probability ~2^-29 per operation)
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def barostat_propose(x, box, group_idxs, volume_scale, rand0, adaptive=True):
    """Proposal of one barostat move in the reference's float32 arithmetic.
    k_setup_barostat_move (k_barostat.cuh:92-115): volume change delta = scale_factor * 2 * (rand0 - 0.5), length scale
    cbrt(V'/V); k_find_group_centroids (:71-88): centroids from float32-rounded coordinates summed in 2^36 fixed point;
    k_rescale_positions (:11-68): centroid scaled about the box centre, group moved rigidly, then shifted so that the
    moved centroid lies in the scaled home box.  Atoms in no group are left where they are.
    Returns (x_proposed, box_proposed, volume, delta_volume, volume_scale_used)."""
    x = np.asarray(x, dtype=np.float64)
    box = np.asarray(box, dtype=np.float64)
    volume = _f32(box[0, 0] * box[1, 1] * box[2, 2])
    scale_factor = float(volume_scale)
    if adaptive and scale_factor == 0.0:
        scale_factor = 0.01 * float(volume)
    delta = _f32((scale_factor * 2) * (float(_f32(rand0)) - 0.5))
    new_volume = _f32(volume + delta)
    scale = _f32(np.cbrt(_f32(new_volume / volume)))
    box_prop = box.copy()
    for c in range(3):
        box_prop[c, c] = box[c, c] * float(scale)
    center = [_f32(box[c, c] * 0.5) for c in range(3)]
    scaled_box = [_f32(box[c, c] * float(scale)) for c in range(3)]
    x_prop = x.copy()
    two36 = np.float32(2.0**36)
    for g in group_idxs:
        g = np.sort(np.asarray(g))
        n = _f32(len(g))
        for c in range(3):
            fixed = int(np.sum(np.rint(x[g, c].astype(np.float32) * two36).astype(np.int64)))
            centroid = _f32(_f32(_f32(fixed) / two36) / n)
            displacement = _f32(_fma32(scale, _f32(centroid - center[c]), center[c]) - centroid)
            moved = _f32(displacement + centroid)
            cells = np.floor(_f32(moved / scaled_box[c]))
            if c < 2:  # the reference binary rounds the product for x and y and fuses it for z (barostat.cu, SASS)
                shift = _f32(displacement - _f32(scaled_box[c] * cells))
            else:
                shift = _fma32(scaled_box[c], -cells, displacement)
            x_prop[g, c] = x[g, c] + float(shift)
    return x_prop, box_prop, volume, delta, scale_factor


def barostat_accepts(u_init_fixed, u_final_fixed, volume, delta_volume, num_molecules, temperature, pressure_bar, rand1):
    """Metropolis test of k_decide_move (k_barostat.cuh:120-189): w = dU + P dV - N kT ln(V'/V); the move is rejected
    iff w > 0 and rand1 > exp(-w / kT).  Energies are 2^36 fixed-point integers (overflowed sums give dU = +inf).
    temperature and pressure pass through float32 like the reference's MonteCarloBarostat<float> members.
    Returns (accepted, w)."""
    kt = BOLTZ * float(_f32(temperature))
    pressure = float(_f32(pressure_bar)) * AVOGADRO * 1e-25
    llmax, llmin = (1 << 63) - 1, -(1 << 63)
    if u_init_fixed >= llmax or u_init_fixed <= llmin or u_final_fixed >= llmax or u_final_fixed <= llmin:
        du = np.float32(np.inf)
    else:
        du = _f32(_f32(int(u_final_fixed) - int(u_init_fixed)) / np.float32(2.0**36))
    new_volume = _f32(_f32(volume) + _f32(delta_volume))
    log_ratio = np.log(_f32(new_volume / _f32(volume)))
    w = _f32(float(du) + pressure * float(delta_volume) - (num_molecules * kt) * float(log_ratio))
    rejected = bool(w > 0 and float(_f32(rand1)) > np.exp(-float(w) / kt))
    return (not rejected), w


def barostat_adapt(volume_scale, attempted, accepted, volume):
    """Adaptive volume-scale rule applied after the counters of an attempt are updated (k_barostat.cuh:157-171).
    Returns (volume_scale, attempted, accepted)."""
    if attempted >= 10:
        if accepted < 0.25 * attempted:
            return volume_scale / 1.1, 0, 0
        if accepted > 0.75 * attempted:
            return min(volume_scale * 1.1, float(volume) * 0.3), 0, 0
    return volume_scale, attempted, accepted


def get_group_indices(bond_list, num_atoms):
    """Connected components of the bond graph, atoms outside every bond as singletons at the end.
    Reference: timemachine/md/barostat/utils.py:43-62 (networkx there; union-find here)."""
    parent = list(range(num_atoms))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    bonded = set()
    for i, j in bond_list:
        bonded.update((int(i), int(j)))
        parent[find(int(i))] = find(int(j))
    comps = {}
    for i in sorted(bonded):
        comps.setdefault(find(i), []).append(i)
    # networkx yields components in order of first appearance of any member in the edge list
    order, seen = [], set()
    for i, j in bond_list:
        for k in (int(i), int(j)):
            r = find(k)
            if r not in seen:
                seen.add(r)
                order.append(r)
    groups = [np.array(comps[r]) for r in order]
    for g in groups:
        assert np.all(np.diff(g) == 1)
    groups += [np.array([i], dtype=np.int32) for i in range(num_atoms) if i not in bonded]
    return groups


# ---------------------------------------------------------------------------------------------------------------------
# Restraints and the precomputed pair list.  Energies restate the reference's Python potentials
# (timemachine/potentials/bonded.py:219-242 flat_bottom_bond; chiral_restraints.py:10-125; nonbonded.py:403-446
# nonbonded_on_precomputed_pairs) and are pinned to them through tests/golden/restraints.npz; gradients are the
# analytic forms of the kernels (k_flat_bottom_bond.cuh:137-170, chiral_utils.cuh:94-181,
# k_nonbonded_precomputed.cuh:88-181), checked against finite differences of the pinned energies.
def flat_bottom_bond(x, params, box, bond_idxs):
    """u = k/4 (r - rmax)^4 for r > rmax, k/4 (r - rmin)^4 for r < rmin, minimum image.  Returns (u, du_dx, du_dp)."""
    x = np.asarray(x, dtype=np.float64)
    params = np.asarray(params, dtype=np.float64).reshape(-1, 3)
    i, j = np.asarray(bond_idxs).reshape(-1, 2).T
    d = delta_r(x[i], x[j], box)
    r = np.linalg.norm(d, axis=1)
    k, rmin, rmax = params.T
    lo, hi = (r < rmin), (r > rmax)
    dlo, dhi = r - rmin, r - rmax
    u = np.sum(k / 4 * (lo * dlo**4 + hi * dhi**4))
    du_dr = k * (lo * dlo**3 + hi * dhi**3)
    g = (du_dr / r)[:, None] * d
    du_dx = np.zeros_like(x)
    np.add.at(du_dx, i, g)
    np.add.at(du_dx, j, -g)
    du_dp = np.stack([(lo * dlo**4 + hi * dhi**4) / 4, lo * (-k * dlo**3), hi * (-k * dhi**3)], axis=1)
    return u, du_dx, du_dp


def centroid_restraint(x, group_a_idxs, group_b_idxs, kb, b0):
    """potentials/bonded.py:8-31: kb (|<x_a> - <x_b>| - b0)^2, geometric centroids, no periodic imaging, no parameters.
    Gradient as k_centroid_restraint.cuh:62-80 (the b0 == 0 form needs no division by the distance).  Returns (u, du_dx)."""
    x = np.asarray(x, dtype=np.float64)
    ga = np.asarray(group_a_idxs).reshape(-1)
    gb = np.asarray(group_b_idxs).reshape(-1)
    delta = x[ga].mean(axis=0) - x[gb].mean(axis=0)
    dij = np.sqrt(np.sum(delta * delta))
    du_dx = np.zeros_like(x)
    if b0 == 0:
        u = kb * dij * dij
        g = 2 * kb * delta
    else:
        u = kb * (dij - b0) ** 2
        g = 2 * kb * (dij - b0) * delta / dij
    np.add.at(du_dx, ga, g / ga.size)
    np.add.at(du_dx, gb, -g / gb.size)
    return u, du_dx


def log_flat_bottom_bond(x, params, box, bond_idxs, beta):
    """u = -sum log(1 - exp(-beta u_fb_b)) / beta over bonds (reference potentials/bonded.py:245-253, kernel
    k_log_flat_bottom_bond.cuh:7-117); gradients are the flat-bottom ones scaled per bond by
    -exp(-beta u_fb) / (1 - exp(-beta u_fb)).  Returns (u, du_dx, du_dp)."""
    x = np.asarray(x, dtype=np.float64)
    params = np.asarray(params, dtype=np.float64).reshape(-1, 3)
    i, j = np.asarray(bond_idxs).reshape(-1, 2).T
    d = delta_r(x[i], x[j], box)
    r = np.linalg.norm(d, axis=1)
    k, rmin, rmax = params.T
    lo, hi = (r < rmin), (r > rmax)
    dlo, dhi = r - rmin, r - rmax
    nrg = k / 4 * (lo * dlo**4 + hi * dhi**4)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.sum(-np.log(-np.expm1(-beta * nrg))) / beta
        e = np.exp(-beta * nrg)
        pre = -e / (1.0 - e)
    du_dr = k * (lo * dlo**3 + hi * dhi**3)
    g = (pre * du_dr / r)[:, None] * d
    du_dx = np.zeros_like(x)
    np.add.at(du_dx, i, g)
    np.add.at(du_dx, j, -g)
    du_dp = pre[:, None] * np.stack([(lo * dlo**4 + hi * dhi**4) / 4, lo * (-k * dlo**3), hi * (-k * dhi**3)], axis=1)
    return u, du_dx, du_dp


def _unit_and_jac(v):
    n = np.linalg.norm(v)
    u = v / n
    return u, (np.eye(3) - np.outer(u, u)) / n


def _cross_jacs(a, b):
    """Rows r: d(a x b)_r / da and / db."""
    ja = np.array([[0, b[2], -b[1]], [-b[2], 0, b[0]], [b[1], -b[0], 0]])
    jb = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return ja, jb


def pyramidal_volume_and_grad(xc, x1, x2, x3):
    """(x^ x y^) . z^ with x = x1 - xc etc.; gradients wrt (xc, x1, x2, x3)  (chiral_utils.cuh:94-137)."""
    xx, yy, zz = x1 - xc, x2 - xc, x3 - xc
    (x, jx), (y, jy), (z, jz) = _unit_and_jac(xx), _unit_and_jac(yy), _unit_and_jac(zz)
    xy = np.cross(x, y)
    ja, jb = _cross_jacs(x, y)
    g1, g2, g3 = (z @ ja) @ jx, (z @ jb) @ jy, xy @ jz
    return float(xy @ z), (-g1 - g2 - g3, g1, g2, g3)


def torsion_volume_and_grad(x0, x1, x2, x3):
    """(x^ x y^) . (y^ x z^) with x = x1 - x0, y = x1 - x2, z = x3 - x2  (chiral_utils.cuh:139-181)."""
    xx, yy, zz = x1 - x0, x1 - x2, x3 - x2
    (x, jx), (y, jy), (z, jz) = _unit_and_jac(xx), _unit_and_jac(yy), _unit_and_jac(zz)
    xy, yz = np.cross(x, y), np.cross(y, z)
    j0x, j0y = _cross_jacs(x, y)
    j1y, j1z = _cross_jacs(y, z)
    gx = (yz @ j0x) @ jx
    gy = (yz @ j0y) @ jy + (xy @ j1y) @ jy
    gz = (xy @ j1z) @ jz
    return float(xy @ yz), (-gx, gx + gy, -gy - gz, gz)


def chiral_atom_restraint(x, params, idxs):
    """u = sum k vol^2 over restraints with vol > 0 (chiral_restraints.py:64-75,103-112).  Returns (u, du_dx, du_dp)."""
    x = np.asarray(x, dtype=np.float64)
    params = np.asarray(params, dtype=np.float64).reshape(-1)
    du_dx, du_dp, u = np.zeros_like(x), np.zeros_like(params), 0.0
    for t, quad in enumerate(np.asarray(idxs).reshape(-1, 4)):
        vol, grads = pyramidal_volume_and_grad(*x[quad])
        if vol > 0:
            u += params[t] * vol * vol
            if params[t] == 0:
                continue  # the kernels skip a restraint with k == 0 entirely, du/dk included (k_chiral_restraint.cuh:60-62)
            du_dp[t] = vol * vol
            for atom, g in zip(quad, grads):
                du_dx[atom] += 2 * params[t] * vol * g
    return u, du_dx, du_dp


def chiral_bond_restraint(x, params, idxs, signs):
    """u = sum k vol^2 over restraints with sign * vol > 0 (chiral_restraints.py:77-100,115-125)."""
    x = np.asarray(x, dtype=np.float64)
    params = np.asarray(params, dtype=np.float64).reshape(-1)
    du_dx, du_dp, u = np.zeros_like(x), np.zeros_like(params), 0.0
    for t, quad in enumerate(np.asarray(idxs).reshape(-1, 4)):
        vol, grads = torsion_volume_and_grad(*x[quad])
        if signs[t] * vol > 0:
            u += params[t] * vol * vol
            if params[t] == 0:
                continue  # as above (k_chiral_restraint.cuh:150-152)
            du_dp[t] = vol * vol
            for atom, g in zip(quad, grads):
                du_dx[atom] += 2 * params[t] * vol * g
    return u, du_dx, du_dp


def nonbonded_precomputed(x, params, box, pair_idxs, beta, cutoff):
    """Pairs with their own (q_ij, sig_ij, eps_ij, w_ij): u = q_ij erfc(beta d) S(d) / d + 4 eps_ij (s^12 - s^6),
    s = sig_ij / d, d^2 = |dx|_pbc^2 + w_ij^2 < cutoff^2 (nonbonded.py:403-446).  Returns (u, du_dx, du_dp[M,4])."""
    x = np.asarray(x, dtype=np.float64)
    params = np.asarray(params, dtype=np.float64).reshape(-1, 4)
    i, j = np.asarray(pair_idxs).reshape(-1, 2).T
    q, sig, eps, w = params.T
    dx = delta_r(x[i], x[j], box)
    d = np.sqrt(np.sum(dx * dx, axis=1) + w * w)
    keep = d < cutoff
    from scipy.special import erfc

    damping = erfc(beta * d) * switch_fn(d)
    d_damping = -2 * beta / np.sqrt(np.pi) * np.exp(-((beta * d) ** 2)) * switch_fn(d) + erfc(beta * d) * d_switch_fn(d)
    u_es = q * damping / d
    du_dd_es = q * (d_damping / d - damping / d**2)
    lj_on = (eps != 0) & (sig != 0)
    s6 = np.where(lj_on, (sig / d) ** 6, 0.0)
    u_lj = 4 * eps * (s6 * s6 - s6)
    du_dd_lj = np.where(lj_on, 24 * eps * (s6 - 2 * s6 * s6) / d, 0.0)
    es_on = q != 0
    u = np.sum(np.where(keep, np.where(es_on, u_es, 0) + u_lj, 0.0))
    du_dd = np.where(keep, np.where(es_on, du_dd_es, 0) + du_dd_lj, 0.0)
    g = (du_dd / d)[:, None] * dx
    du_dx = np.zeros_like(x)
    np.add.at(du_dx, i, g)
    np.add.at(du_dx, j, -g)
    du_dp = np.zeros_like(params)
    du_dp[:, 0] = np.where(keep & es_on, damping / d, 0)
    du_dp[:, 1] = np.where(keep & lj_on, 4 * eps * (12 * s6 * s6 - 6 * s6) / np.where(sig != 0, sig, 1.0), 0)
    du_dp[:, 2] = np.where(keep & lj_on, 4 * (s6 * s6 - s6), 0)
    du_dp[:, 3] = (du_dd / d) * w
    return u, du_dx, du_dp


# ---------------------------------------------------------------------------------------------------------------------
# water exchange by biased deletion (timemachine/md/exchange/exchange_mover.py:64-234) - test oracle for
# timemachine_b200/csrc/exchange.cu.  PINNING: tests/golden/exchange.npz holds outputs of the reference's own Python
# (make_golden_exchange.py executes nonbonded_block_unsummed per molecule as BDExchangeMove.batch_log_weights does, and
# get_water_groups / compute_raw_ratio_given_weights / delta_r_np from exchange_mover.py); tests/test_oracle_exchange.py
# holds the functions below to it at 1e-11.  The rotation is checked against the Hamilton-product definition, and the
# whole movers are cross-checked on the GPU against the compiled reference (tests/test_exchange_gpu.py).
def pair_energy_matrix(x, params, box, rows, cols, beta, cutoff):
    """u_ij for i in rows, j in cols (0 outside the cutoff, NaN on a coincident pair), like
    nonbonded_block_unsummed (nonbonded.py:82-150) / k_atom_by_atom_energies (k_nonbonded.cuh:604-700)."""
    rows, cols = np.asarray(rows), np.asarray(cols)
    xi, xj, pi, pj = x[rows], x[cols], params[rows], params[cols]
    dxyz = delta_r(xi[:, None, :], xj[None, :, :], box)
    dw = pi[:, None, 3] - pj[None, :, 3]
    d2 = np.sum(dxyz * dxyz, axis=-1) + dw * dw
    keep = d2 < cutoff * cutoff
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.sqrt(d2)
        t = _pair_terms(d, pi[:, None, 0], pj[None, :, 0], pi[:, None, 1], pj[None, :, 1], pi[:, None, 2], pj[None, :, 2], beta)
    return np.where(keep, t["u"], 0.0)


def mol_energies(x, params, box, mols, beta, cutoff):
    """Energy of every molecule with all atoms outside it (exchange_mover.py:100-138 batch_U_fn)."""
    n = x.shape[0]
    out = []
    for m in mols:
        others = np.delete(np.arange(n), np.asarray(m))
        out.append(float(np.sum(pair_energy_matrix(x, params, box, m, others, beta, cutoff))))
    return np.array(out)


def bd_log_weights(x, params, box, mols, nb_beta, cutoff, temperature):
    """exchange_mover.py:140-152 batch_log_weights: beta * U_mol."""
    return mol_energies(x, params, box, mols, nb_beta, cutoff) / (BOLTZ * temperature)


def logsumexp(v):
    v = np.asarray(v, dtype=np.float64)
    m = np.max(v)
    return float(m + np.log(np.sum(np.exp(v - m))))


def bd_log_acceptance(log_weights_before, log_weights_after):
    """exchange_mover.py:229: min(logsumexp(before) - logsumexp(after), 0)."""
    return min(logsumexp(log_weights_before) - logsumexp(log_weights_after), 0.0)


def quaternion_rotate(coords, q):
    """Rotate points by the (normalised) quaternion q = (w, x, y, z): q (0, v) q*  (k_rotations.cu:9-48)."""
    w, x, y, z = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    R = np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )
    return np.asarray(coords) @ R.T


def rotate_and_translate_mol(coords, box, q, translation, scale=True):
    """Rotate about the centroid, then put the centroid at the translation imaged into the home box
    (exchange_mover.py:29-42 randomly_rotate_and_translate with a given rotation; k_rotations.cu:112-193)."""
    bd = np.diag(box)
    t = np.asarray(translation, dtype=np.float64) * (bd if scale else 1.0)
    t = t - bd * np.floor(t / bd)
    c = np.mean(coords, axis=0, keepdims=True)
    return quaternion_rotate(coords - c, q) + t


def water_groups(coords, box, center, mols, radius):
    """exchange_mover.py:270-281 get_water_groups: molecules whose centroid lies within `radius` of `center` (minimum
    image) and the rest."""
    c = np.array([np.mean(coords[np.asarray(m)], axis=0) for m in mols])
    d = np.linalg.norm(delta_r(c, np.asarray(center), box), axis=1)
    return np.nonzero(d < radius)[0], np.nonzero(d >= radius)[0]


def _proposal_probability(n_a, n_b):
    """exchange_mover.py:283-296."""
    assert n_a >= 0 and n_b >= 0 and (n_a > 0 or n_b > 0)
    return 0.5 if (n_a > 0 and n_b > 0) else 1.0


def tibd_raw_log_probability(log_weights_src_before, log_weights_dest_after, n_src, n_dest, vol_src, vol_dest):
    """exchange_mover.py:298-323 compute_raw_ratio_given_weights: a molecule leaves a region of n_src molecules and
    volume vol_src for a region of n_dest molecules and volume vol_dest; log_weights_dest_after includes the moved one."""
    g_fwd = _proposal_probability(n_src, n_dest)
    g_rev = _proposal_probability(n_src - 1, n_dest + 1)
    return (
        logsumexp(log_weights_src_before) - logsumexp(log_weights_dest_after) + np.log(vol_dest) - np.log(vol_src)
        + np.log(g_rev) - np.log(g_fwd)
    )
