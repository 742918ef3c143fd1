"""Generate tests/golden/hrex_driver.npz: the reference's own HREX host functions run on seeded inputs.

Run here (the container that has /root/reference), never on the GPU box:

    python tests/golden/make_golden_hrex.py

The functions are exec'd UNMODIFIED from the reference source text (timemachine/md/hrex.py and
timemachine/fe/free_energy.py import jax / rdkit at module scope, which this image does not have; the functions below
are pure numpy + scipy):
  * md/hrex.py: get_normalized_kl_divergence, get_cumulative_replica_state_counts, estimate_transition_matrix,
    estimate_relaxation_time, get_samples_by_iter_by_replica, HREXDiagnostics.cumulative_swap_acceptance_rates,
    get_swap_attempts_per_iter_heuristic
  * fe/free_energy.py: compute_potential_matrix (with a stand-in potential whose execute_batch(_sparse) returns a known
    function of (coords index, params index): this pins the index bookkeeping), verify_and_sanitize_potential_matrix
"""

from __future__ import annotations

import types
import warnings
from pathlib import Path

import numpy as np
from scipy.stats import entropy

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def between(src: str, start: str, end: str) -> str:
    i = src.index(start)
    return src[i : src.index(end, i)]


def main():
    rng = np.random.default_rng(2024)
    hrex_src = (REF / "timemachine/md/hrex.py").read_text()
    fe_src = (REF / "timemachine/fe/free_energy.py").read_text()

    ns = {
        "np": np, "entropy": entropy, "Sequence": list, "ReplicaIdx": int, "NDArray": np.ndarray, "Samples": object,
        "not_ragged": lambda xss: len({len(xs) for xs in xss}) <= 1,
    }
    exec(between(hrex_src, "def get_normalized_kl_divergence", "@dataclass\nclass HREXDiagnostics"), ns)
    exec(between(hrex_src, "def get_swap_attempts_per_iter_heuristic", "def run_hrex("), ns)

    # a plausible permutation history: random neighbour transpositions
    n_states, n_iters = 6, 60
    perm = np.arange(n_states)
    history = []
    for _ in range(n_iters):
        for _ in range(3):
            s = rng.integers(0, n_states - 1)
            if rng.random() < 0.6:
                perm[s], perm[s + 1] = perm[s + 1], perm[s]
        history.append(perm.copy())
    history = np.array(history)
    counts = ns["get_cumulative_replica_state_counts"](history)
    tm = ns["estimate_transition_matrix"](history)
    relax = ns["estimate_relaxation_time"](tm)
    kl = ns["get_normalized_kl_divergence"](history)
    samples = [[f"i{it}s{s}" for s in range(n_states)] for it in range(n_iters)]
    by_replica = ns["get_samples_by_iter_by_replica"](samples, history.tolist())
    by_replica_codes = np.array([[int(x[1:].split("s")[0]) * 100 + int(x.split("s")[1]) for x in xs] for xs in by_replica])
    frac = rng.integers(1, 20, size=(n_iters, n_states - 1, 2))
    frac[..., 0] = np.minimum(frac[..., 0], frac[..., 1])  # accepted <= proposed
    n_accepted, n_proposed = np.moveaxis(np.array(frac), -1, 0)  # HREXDiagnostics.cumulative_swap_acceptance_rates body
    cum_rates = np.cumsum(n_accepted, axis=0) / np.cumsum(n_proposed, axis=0)
    assert "np.cumsum(n_accepted, axis=0) / np.cumsum(n_proposed, axis=0)" in hrex_src

    # ---- compute_potential_matrix / verify_and_sanitize_potential_matrix --------------------------------------------
    class IndeterminateEnergyWarning(UserWarning):
        pass

    ns2 = {
        "np": np, "NDArray": np.ndarray, "Optional": __import__("typing").Optional, "Sequence": list,
        "custom_ops": types.SimpleNamespace(Potential=object), "HREX": dict, "CoordsVelBox": object,
        "warn": warnings.warn, "IndeterminateEnergyWarning": IndeterminateEnergyWarning,
    }
    exec(between(fe_src, "def compute_potential_matrix(", "def make_u_kl_fxn("), ns2)

    n = 7
    coords = rng.normal(size=(n, 5, 3))
    boxes = np.stack([np.eye(3) * (3 + i) for i in range(n)])
    params = rng.normal(size=(n, 4))

    def energy(ci, pi):  # any function that identifies the (coords, params) pair it was called with
        return coords[ci].sum() * 10 + params[pi].sum() + boxes[ci][0, 0]

    class FakePotential:
        def execute_batch_sparse(self, cs, ps, bs, cidx, pidx, dx, dp, du):
            assert (dx, dp, du) == (False, False, True)
            return None, None, np.array([energy(c, p) for c, p in zip(cidx, pidx)])

        def execute_batch(self, cs, ps, bs, dx, dp, du):
            return None, None, np.array([[energy(c, p) for p in range(len(ps))] for c in range(len(cs))])

    replica_idx_by_state = rng.permutation(n)
    hrex = types.SimpleNamespace(
        replicas=[types.SimpleNamespace(coords=coords[i], box=boxes[i]) for i in range(n)],
        replica_idx_by_state=replica_idx_by_state.tolist(),
    )
    out = {}
    for k in (1, 2, None):
        out[f"U_kl_k{k}"] = ns2["compute_potential_matrix"](FakePotential(), hrex, params, k)
    dirty = out["U_kl_k2"].copy()
    off = [(r, s) for r in range(n) for s in range(n) if np.isfinite(dirty[r, s]) and replica_idx_by_state[s] != r]
    dirty[off[0]] = np.nan
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clean = ns2["verify_and_sanitize_potential_matrix"](dirty, replica_idx_by_state.tolist())

    np.savez(
        OUT / "hrex_driver.npz", history=history, counts=counts, transition_matrix=tm, relaxation_time=relax, kl=kl,
        by_replica_codes=by_replica_codes, frac=frac, cum_rates=cum_rates,
        swap_heuristic=np.array([ns["get_swap_attempts_per_iter_heuristic"](k) for k in range(1, 10)]),
        coords=coords, boxes=boxes, params=params, replica_idx_by_state=replica_idx_by_state, dirty=dirty, clean=clean, **out,
    )
    print(f"hrex_driver: relaxation_time={relax:.6f} kl={kl:.6f} finite entries k=1: {np.isfinite(out['U_kl_k1']).sum()}")


if __name__ == "__main__":
    main()
