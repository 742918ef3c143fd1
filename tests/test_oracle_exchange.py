"""CPU: the exchange-move part of the NumPy oracle (oracle/tm_oracle.py, "water exchange by biased deletion").

The per-molecule energies are tied to the interaction-group energy of the same oracle, which is pinned to the reference's
own `nonbonded` through tests/golden/nonbonded_*.npz (tests/test_oracle.py); the quaternion rotation is checked against the
Hamilton-product definition the reference's kernel uses (k_rotations.cu:9-48)."""

import numpy as np

from oracle import tm_oracle as O
from tests.common import water_box


def hamilton(q1, q2):
    w1, x1, y1, z1 = q1
    w2, x2, y2, z2 = q2
    return np.array(
        [
            w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
            w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
            w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
        ]
    )


def test_quaternion_rotation_is_q_v_qconj():
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.normal(size=4) * rng.uniform(0.1, 5.0)
        v = rng.normal(size=(4, 3))
        qn = q / np.linalg.norm(q)
        conj = qn * np.array([1, -1, -1, -1])
        ref = np.array([hamilton(hamilton(qn, np.concatenate([[0.0], p])), conj)[1:] for p in v])
        got = O.quaternion_rotate(v, q)
        np.testing.assert_allclose(got, ref, atol=1e-13)
        np.testing.assert_allclose(np.linalg.norm(got, axis=1), np.linalg.norm(v, axis=1), rtol=1e-13)


def test_rotate_and_translate_puts_the_centroid_at_the_imaged_translation():
    rng = np.random.default_rng(1)
    box = np.diag([2.0, 3.0, 4.0])
    mol = rng.normal(size=(3, 3))
    for _ in range(10):
        t = rng.uniform(-2, 3, 3)
        out = O.rotate_and_translate_mol(mol, box, rng.normal(size=4), t)
        want = (t * np.diag(box)) % np.diag(box)
        np.testing.assert_allclose(out.mean(0), want, atol=1e-12)
        d0 = np.linalg.norm(mol[0] - mol[1])
        np.testing.assert_allclose(np.linalg.norm(out[0] - out[1]), d0, rtol=1e-12)


def test_mol_energies_are_interaction_group_energies_and_incremental_weights_are_consistent():
    s = water_box(40, seed=3)
    x, params, box = s["x"], s["params"], s["box"]
    n = len(x)
    mols = [[3 * i, 3 * i + 1, 3 * i + 2] for i in range(40)]
    e = O.mol_energies(x, params, box, mols, 2.0, 1.2)
    for m in (0, 17, 39):
        others = np.delete(np.arange(n), mols[m])
        u, _, _ = O.nonbonded_interaction_group(x, params, box, mols[m], others, 2.0, 1.2)
        np.testing.assert_allclose(e[m], u, rtol=1e-12)
    # every water-water pair is counted in exactly two molecule energies
    u_all, _, _ = O.nonbonded_all_pairs(x, params, box, 2.0, 1.2)
    intra = sum(float(np.sum(np.triu(O.pair_energy_matrix(x, params, box, m, m, 2.0, 1.2), 1))) for m in mols)
    np.testing.assert_allclose(e.sum(), 2.0 * (u_all - intra), rtol=1e-10)
    # moving one molecule changes the others' weights only through their pair energies with it (the transposition
    # trick of exchange_mover.py:156-199)
    w0 = O.bd_log_weights(x, params, box, mols, 2.0, 1.2, 300.0)
    moved = x.copy()
    moved[mols[5]] = O.rotate_and_translate_mol(x[mols[5]], box, [0.3, -1.0, 0.2, 0.9], [0.1, 0.7, 0.4])
    w1 = O.bd_log_weights(moved, params, box, mols, 2.0, 1.2, 300.0)
    beta = 1.0 / (O.BOLTZ * 300.0)
    for m in (0, 6, 20):
        old = np.sum(O.pair_energy_matrix(x, params, box, mols[5], mols[m], 2.0, 1.2))
        new = np.sum(O.pair_energy_matrix(moved, params, box, mols[5], mols[m], 2.0, 1.2))
        np.testing.assert_allclose(w1[m], w0[m] + beta * (new - old), rtol=1e-9, atol=1e-9)
    assert O.bd_log_acceptance(w0, w0) == 0.0
    np.testing.assert_allclose(O.logsumexp([1000.0, 1000.0]), 1000.0 + np.log(2.0))


def test_exchange_oracle_matches_the_references_own_python():
    """tests/golden/exchange.npz was written by executing the reference's Python (make_golden_exchange.py):
    nonbonded_block_unsummed per molecule as BDExchangeMove.batch_log_weights uses it, get_water_groups,
    compute_raw_ratio_given_weights, delta_r_np.  The oracle's restatements must reproduce it."""
    from pathlib import Path

    g = np.load(Path(__file__).resolve().parent / "golden" / "exchange.npz")
    x, params, box, mols = g["x"], g["params"], g["box"], g["mols"]
    beta, cutoff, temperature = float(g["beta"]), float(g["cutoff"]), float(g["temperature"])
    np.testing.assert_allclose(O.mol_energies(x, params, box, mols, beta, cutoff), g["mol_energy"], rtol=1e-11, atol=1e-10)
    np.testing.assert_allclose(O.bd_log_weights(x, params, box, mols, beta, cutoff, temperature), g["log_weights"], rtol=1e-11, atol=1e-10)
    others = np.delete(np.arange(len(x)), mols[4])
    np.testing.assert_allclose(O.pair_energy_matrix(x, params, box, mols[4], others, beta, cutoff), g["pair_rows_mol4"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(O.delta_r(x[:10], x[10:20], box), g["delta_r"], rtol=0, atol=1e-15)
    inner, outer = O.water_groups(x, box, g["center"], mols, float(g["radius"]))
    np.testing.assert_array_equal(inner, g["inner"])
    np.testing.assert_array_equal(outer, g["outer"])
    lw, after, moved = g["log_weights"], g["after"], int(g["moved"])
    vi, vo = float(g["vol_inner"]), float(g["vol_outer"])
    raw_in = O.tibd_raw_log_probability(lw[outer], after[np.append(inner, moved)], len(outer), len(inner), vo, vi)
    raw_out = O.tibd_raw_log_probability(lw[inner], after[np.append(outer, inner[0])], len(inner), len(outer), vi, vo)
    np.testing.assert_allclose([raw_in, raw_out], [g["raw_in"], g["raw_out"]], rtol=1e-13)
    np.testing.assert_allclose(O.tibd_raw_log_probability(lw[:5], after[:1], 5, 0, 2.0, 3.0), g["raw_empty_dest"], rtol=1e-13)
    np.testing.assert_allclose(O.tibd_raw_log_probability(lw[:1], after[:4], 1, 3, 2.0, 3.0), g["raw_single_src"], rtol=1e-13)
