"""Host-side logic of the one-replica-per-GPU layout.

The reference runs HREX replicas sequentially through ONE Context on ONE GPU (fe/free_energy.py:1383-1618) and farms
independent windows out as separate processes (parallel/client.py:188-218); it has no collective anywhere.  Here every
rank owns one replica (its own Context: coordinates, velocities, box, neighbour list, RNG stream) and holds all K
parameter sets.  Per frame each rank evaluates its replica's energy under the neighbouring windows' parameters, the
K-vector rows are all-gathered (K doubles per rank: latency-bound, off the per-step path), and every rank runs the
same seeded neighbour-swap sweep, so all ranks reach the same permutation with no further traffic.  A swap re-binds
PARAMETERS on the device (BoundPotential.set_params_device); coordinates never cross NVLink.

Backend-agnostic: works with torch.distributed over NCCL (GPU) or gloo (the CPU tests).
"""

from __future__ import annotations

import numpy as np

BOLTZ = 0.008314462618  # kJ/mol/K (reference timemachine/cpp/src/constants.hpp:5)


def candidate_states(state: int, n_states: int, max_delta: int = 1) -> list[int]:
    """States whose energies replica-in-`state` must evaluate (cf. `compute_sparse`, fe/free_energy.py:1177-1192)."""
    return [k for k in range(state - max_delta, state + max_delta + 1) if 0 <= k < n_states]


def energy_row(n_states: int, states: list[int], energies: list[float]) -> np.ndarray:
    """Length-K row of reduced-unit-free energies (kJ/mol); +inf where not evaluated."""
    row = np.full(n_states, np.inf, dtype=np.float64)
    for k, u in zip(states, energies):
        row[k] = u
    return row


def all_gather_rows(row: np.ndarray, dist=None, device=None) -> np.ndarray:
    """All-gather one length-K float64 row per rank into the K x K matrix u[replica, state]."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return row.reshape(1, -1).copy()
    import torch

    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(row, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * row.size, dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().reshape(world, row.size)


def neighbour_swaps(u_matrix: np.ndarray, states: np.ndarray, temperature: float, rng: np.random.Generator) -> np.ndarray:
    """One deterministic sweep of neighbour swap attempts (reference md/hrex.py:132-240).

    u_matrix[r, s]: energy of replica r's coordinates under state s (inf = not evaluated);  states[r]: state currently
    held by replica r.  Metropolis on  exp(-(u_a(s+1) + u_b(s) - u_a(s) - u_b(s+1)) / kT).  Every rank calls this with
    the same matrix and an identically seeded generator, so all ranks agree.
    """
    kT = BOLTZ * temperature
    states = np.array(states, copy=True)
    K = len(states)
    replica_of_state = np.argsort(states)
    for s in range(K - 1):
        ra, rb = int(replica_of_state[s]), int(replica_of_state[s + 1])
        cur = u_matrix[ra, s] + u_matrix[rb, s + 1]
        new = u_matrix[ra, s + 1] + u_matrix[rb, s]
        draw = rng.random()  # always consume one number so the stream stays aligned across ranks
        if not (np.isfinite(cur) and np.isfinite(new)):
            continue
        if np.log(draw) < -(new - cur) / kT:
            states[ra], states[rb] = s + 1, s
            replica_of_state[s], replica_of_state[s + 1] = rb, ra
    return states


def run_neighbor_swaps(replica_idx_by_state, neighbor_pairs, log_q_kl, pair_idxs, uniform_samples):
    """A batch of neighbour-swap attempts with the reference's semantics (timemachine/md/hrex.py:50-129,
    `_run_neighbor_swaps`): attempt i picks the state pair neighbor_pairs[pair_idxs[i]] = (s_a, s_b), held by replicas
    (r_a, r_b); it is accepted iff uniform_samples[i] < exp(min(0, log_q[r_a, s_b] + log_q[r_b, s_a] - log_q[r_a, s_a]
    - log_q[r_b, s_b])), and then the two states exchange replicas.  log_q_kl[r, s] = -u(x_r; state s) / kT.
    Deterministic given (pair_idxs, uniform_samples): every rank runs it on the all-gathered matrix and agrees.
    Returns (replica_idx_by_state, proposed_by_pair, accepted_by_pair)."""
    replica_idx_by_state = np.array(replica_idx_by_state, copy=True)
    neighbor_pairs = np.asarray(neighbor_pairs)
    log_q_kl = np.asarray(log_q_kl, dtype=np.float64)
    proposed = np.zeros(len(neighbor_pairs), dtype=np.uint32)
    accepted = np.zeros(len(neighbor_pairs), dtype=np.uint32)
    for pair_idx, u in zip(np.asarray(pair_idxs), np.asarray(uniform_samples)):
        s_a, s_b = neighbor_pairs[pair_idx]
        proposed[pair_idx] += 1
        r_a, r_b = replica_idx_by_state[s_a], replica_idx_by_state[s_b]
        log_q_diff = (log_q_kl[r_a, s_b] + log_q_kl[r_b, s_a]) - (log_q_kl[r_a, s_a] + log_q_kl[r_b, s_b])
        if u < np.exp(min(log_q_diff, 0.0)):
            replica_idx_by_state[s_a], replica_idx_by_state[s_b] = r_b, r_a
            accepted[pair_idx] += 1
    return replica_idx_by_state, proposed, accepted


def run_neighbor_swaps_native(replica_idx_by_state, neighbor_pairs, log_q_kl, pair_idxs, uniform_samples):
    """`run_neighbor_swaps` through the C ABI (`tmb_hrex_run_neighbor_swaps`, host code of libtmb200.so): the reference
    jit-compiles this loop (md/hrex.py:50); n_states^3 attempts per frame in Python would cost milliseconds of idle GPU on
    every rank.  Same results as the Python statement above, bit for bit (tests/test_replica_cpu.py)."""
    import ctypes as C

    from ._lib import load

    L = load()
    perm = np.ascontiguousarray(replica_idx_by_state, dtype=np.int32).copy()
    pairs = np.ascontiguousarray(neighbor_pairs, dtype=np.int32).reshape(-1, 2)
    q = np.ascontiguousarray(log_q_kl, dtype=np.float64)
    pidx = np.ascontiguousarray(pair_idxs, dtype=np.int32)
    us = np.ascontiguousarray(uniform_samples, dtype=np.float64)
    assert q.ndim == 2 and q.shape[1] == len(perm) and len(pidx) == len(us)
    proposed = np.zeros(len(pairs), dtype=np.uint32)
    accepted = np.zeros(len(pairs), dtype=np.uint32)

    def ptr(a, t):
        return a.ctypes.data_as(C.POINTER(t))

    rc = L.tmb_hrex_run_neighbor_swaps(
        len(perm), q.shape[0], len(pairs), ptr(pairs, C.c_int32), ptr(q, C.c_double), len(pidx), ptr(pidx, C.c_int32),
        ptr(us, C.c_double), ptr(perm, C.c_int32), ptr(proposed, C.c_uint32), ptr(accepted, C.c_uint32),
    )
    if rc != 0:
        raise RuntimeError(L.tmb_last_error().decode())
    return perm, proposed, accepted


def i128_to_energy(lo: int, hi: int) -> float:
    """Fixed-point int128 energy -> kJ/mol, NaN when it left the int64 range (reference wrap_kernels.cpp:83-89)."""
    v = (int(hi) << 64) | (int(lo) & ((1 << 64) - 1))
    if v >= (1 << 63) - 1 or v <= -(1 << 63):
        return float("nan")
    return v / float(1 << 36)
