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
