"""CPU: the restraint / precomputed-pair oracles (oracle/tm_oracle.py) against golden values computed by
the reference's own Python potentials (tests/golden/make_golden.py -> restraints.npz): energies to 1e-10, analytic
gradients against central finite differences of the reference energies."""

from pathlib import Path

import numpy as np
import pytest

from oracle import tm_oracle as O

G = dict(np.load(Path(__file__).parent / "golden" / "restraints.npz"))


def test_chiral_atom_restraint_oracle():
    u, du_dx, du_dp = O.chiral_atom_restraint(G["x"], G["k_atom"], G["quads"])
    np.testing.assert_allclose(u, G["u_chiral_atom"], rtol=1e-10)
    np.testing.assert_allclose(du_dx, G["chiral_atom_du_dx_fd"], rtol=2e-5, atol=2e-5)
    assert np.all(du_dp >= 0) and np.count_nonzero(du_dp) >= 5  # vol^2 where the restraint is active


def test_chiral_bond_restraint_oracle():
    u, du_dx, du_dp = O.chiral_bond_restraint(G["x"], G["k_bond"], G["quads"], G["signs"])
    np.testing.assert_allclose(u, G["u_chiral_bond"], rtol=1e-10)
    np.testing.assert_allclose(du_dx, G["chiral_bond_du_dx_fd"], rtol=2e-5, atol=2e-5)
    # flipping every sign activates exactly the complementary set of restraints
    u_flip, _, du_dp_flip = O.chiral_bond_restraint(G["x"], G["k_bond"], G["quads"], -G["signs"])
    assert not np.any((du_dp > 0) & (du_dp_flip > 0))
    vols2 = du_dp + du_dp_flip
    np.testing.assert_allclose(u + u_flip, np.sum(G["k_bond"] * vols2), rtol=1e-12)


def test_flat_bottom_bond_oracle():
    u, du_dx, du_dp = O.flat_bottom_bond(G["x"], G["fb_params"], G["box"], G["fb_idxs"])
    np.testing.assert_allclose(u, G["u_fb"], rtol=1e-10)
    np.testing.assert_allclose(du_dx, G["fb_du_dx_fd"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(du_dp, G["fb_du_dp_fd"], rtol=2e-5, atol=2e-5)


def test_nonbonded_precomputed_oracle():
    u, du_dx, du_dp = O.nonbonded_precomputed(G["x"], G["pre_params"], G["box"], G["pre_idxs"], float(G["beta"]), float(G["cutoff"]))
    np.testing.assert_allclose(u, G["u_pre"], rtol=1e-10)
    np.testing.assert_allclose(du_dx, G["pre_du_dx_fd"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(du_dp, G["pre_du_dp_fd"], rtol=5e-5, atol=5e-5)


def test_volume_gradients_against_finite_differences():
    rng = np.random.default_rng(3)
    for fn in (O.pyramidal_volume_and_grad, O.torsion_volume_and_grad):
        pts = rng.normal(size=(4, 3))
        vol, grads = fn(*pts)
        assert -1.0 <= vol <= 1.0
        for a in range(4):
            for c in range(3):
                hp, hm = pts.copy(), pts.copy()
                hp[a, c] += 1e-6
                hm[a, c] -= 1e-6
                fd = (fn(*hp)[0] - fn(*hm)[0]) / 2e-6
                assert grads[a][c] == pytest.approx(fd, abs=1e-7)
        np.testing.assert_allclose(np.sum(grads, axis=0), 0, atol=1e-12)  # translation invariance


def test_log_flat_bottom_bond_matches_reference_python():
    """oracle.log_flat_bottom_bond against the reference's `log_flat_bottom_bond` (bonded.py:245-253), executed by
    tests/golden/make_golden_logfb.py; gradients against finite differences of the reference energy."""
    from pathlib import Path

    import numpy as np

    from oracle import tm_oracle as O

    g = np.load(Path(__file__).parent / "golden" / "log_flat_bottom_bond.npz")
    u, du_dx, du_dp = O.log_flat_bottom_bond(g["x"], g["params"], g["box"], g["idxs"], float(g["beta"]))
    np.testing.assert_allclose(u, g["u"], rtol=1e-12)
    np.testing.assert_allclose(du_dx, g["du_dx_fd"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(du_dp, g["du_dp_fd"], rtol=2e-5, atol=1e-6)
