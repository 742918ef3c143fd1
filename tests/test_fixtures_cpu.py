"""CPU: fixtures and small oracle helpers added in round 2."""

from fractions import Fraction

import numpy as np


def test_dhfr_fixture_is_the_reference_benchmark_geometry():
    """tests/golden/dhfr_5dfr.npz holds the atoms of timemachine/testsystems/data/5dfr_solv_equil.pdb (BASELINE configs[1]:
    23,558 atoms, 62.23 A cubic box, tests/test_benchmark.py:506-518); tests/dhfr_system.py derives a topology from it."""
    from tests.dhfr_system import load_dhfr

    s = load_dhfr()
    assert s["N"] == 23558 and s["n_protein"] == 2489
    np.testing.assert_allclose(np.diag(s["box"]), 6.223)
    n_waters = (s["N"] - s["n_protein"]) // 3
    assert n_waters == 7023
    # every atom is bonded, hydrogens exactly once
    deg = np.bincount(s["bond_idxs"].reshape(-1), minlength=s["N"])
    assert deg.min() >= 1 and deg.max() <= 4
    assert len(s["angle_idxs"]) > n_waters and len(s["proper_idxs"]) > 5000 and len(s["improper_idxs"]) > 400
    # neutral, with exclusions for every bond and angle
    assert abs(s["params"][:, 0].sum()) < 1e-9
    excl = {tuple(e) for e in s["exclusion_idxs"].tolist()}
    assert all((min(a, b), max(a, b)) in excl for a, b in s["bond_idxs"].tolist())
    assert s["scale_factors"].min() > 0 and s["scale_factors"].max() == 1.0
    # hydrogen mass repartitioning conserved the total mass
    assert s["masses"].min() > 3.0
    # the coordinates are the PDB's (three decimals in angstrom)
    np.testing.assert_allclose(s["x"] * 1e4, np.rint(s["x"] * 1e4), atol=1e-6)


def test_fma32_exact_is_a_correctly_rounded_float32_fma():
    from oracle.tm_oracle import fma32_exact

    rng = np.random.default_rng(0)
    a = rng.normal(size=4000).astype(np.float32)
    b = rng.normal(size=4000).astype(np.float32)
    c = (rng.normal(size=4000) * 10.0 ** rng.integers(-6, 3, 4000)).astype(np.float32)
    # force exact float32 midpoints with a non-zero residual: a*b is a midpoint of two floats, c far below one ulp
    a[:200] = np.float32(1.0) + np.float32(2.0**-12)
    b[:200] = np.float32(1.0) + np.float32(2.0**-12)  # (1 + 2^-12)^2 = 1 + 2^-11 + 2^-24: the midpoint of 1 + 2^-11 and the next float
    c[:100] = np.float32(2.0**-60)
    c[100:200] = np.float32(-(2.0**-60))
    r = fma32_exact(a, b, c)
    for i in list(range(300)) + list(range(300, 4000, 37)):
        t = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
        x = np.float32(float(t))
        cands = [np.nextafter(x, np.float32(-np.inf)), x, np.nextafter(x, np.float32(np.inf))]
        best = min(cands, key=lambda y: (abs(Fraction(float(y)) - t), int(np.float32(y).view(np.uint32)) & 1))
        assert best == r[i], (i, a[i], b[i], c[i])
    assert r[0] != r[100]  # the residual's sign decided the two tie cases differently
