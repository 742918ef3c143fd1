"""CPU: the C/OpenMP oracle (oracle/tm_oracle_c.c, used for large sizes and the CPU baseline) agrees with the NumPy
oracle, which is itself pinned to the reference's Python functions."""

import numpy as np

from oracle import build_oracle as OC
from oracle import tm_oracle as O
from tests.common import random_nonbonded_system, water_box


def test_c_nonbonded_matches_numpy_oracle():
    x, params, box = random_nonbonded_system(400, seed=1, w_pattern="some")
    idx = np.arange(400)
    u, dx, dp = OC.nonbonded_block(x, params, box, idx, idx, 2.0, 1.2, True, True, True)
    ru, rdx, rdp = O.nonbonded_all_pairs(x, params, box, 2.0, 1.2)
    np.testing.assert_allclose(u, ru, rtol=1e-11)
    np.testing.assert_allclose(dx, rdx, rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(dp, rdp, rtol=1e-9, atol=1e-8)
    rows, cols = np.arange(0, 40), np.arange(40, 400)
    u, dx, dp = OC.nonbonded_block(x, params, box, rows, cols, 2.0, 1.2, False, True, True)
    ru, rdx, rdp = O.nonbonded_interaction_group(x, params, box, rows, cols, 2.0, 1.2)
    np.testing.assert_allclose(u, ru, rtol=1e-11)
    np.testing.assert_allclose(dx, rdx, rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(dp, rdp, rtol=1e-9, atol=1e-8)


def test_c_water_system_matches_numpy_oracle():
    s = water_box(120, seed=2)
    N = s["N"]
    dx = np.zeros((N, 3))
    idx = np.arange(N)
    u, dxa, _ = OC.nonbonded_block(s["x"], s["params"], s["box"], idx, idx, 2.0, 1.2, True)
    dx += dxa
    u += OC.nonbonded_pairs(s["x"], s["params"], s["box"], s["exclusion_idxs"], s["scale_factors"], -1.0, 2.0, 1.2, dx=dx)
    u += OC.harmonic_bond(s["x"], s["bond_params"], s["bond_idxs"], dx)
    u += OC.harmonic_angle(s["x"], s["angle_params"], s["angle_idxs"], dx)
    ru, rdx, _ = O.nonbonded(s["x"], s["params"], s["box"], s["exclusion_idxs"], s["scale_factors"], 2.0, 1.2)
    ub, dxb, _ = O.harmonic_bond(s["x"], s["bond_params"], s["bond_idxs"])
    ua, dxa2, _ = O.harmonic_angle(s["x"], s["angle_params"], s["angle_idxs"])
    np.testing.assert_allclose(u, ru + ub + ua, rtol=1e-10)
    np.testing.assert_allclose(dx, rdx + dxb + dxa2, rtol=1e-8, atol=1e-7)


def test_c_baoab_matches_numpy(rng):
    n = 50
    x, v = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    du = rng.normal(0, 50, (n, 3))
    masses = rng.uniform(1, 16, n)
    ca, cb, cc = O.langevin_coefficients(300.0, 2.5e-3, 1.0, masses)
    noise = rng.normal(size=(n, 3))
    rx, rv = O.baoab_step(x, v, -du, ca, cb, cc, 2.5e-3, noise)
    x2, v2 = x.copy(), v.copy()
    OC.baoab(x2, v2, du, ca, cb, cc, 2.5e-3, noise)
    np.testing.assert_allclose(x2, rx, rtol=1e-14)
    np.testing.assert_allclose(v2, rv, rtol=1e-14)
