"""CPU: pin oracle/tm_oracle.py to the golden vectors produced by the reference's own Python functions
(tests/golden/make_golden.py) and check the internal consistency of its analytic gradients."""

import numpy as np
import pytest

from oracle import tm_oracle as O


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_nonbonded_energy_matches_reference_python(golden, tag):
    g = golden(f"nonbonded_{tag}")
    args = (g["x"], g["params"], g["box"], g["exclusion_idxs"], g["scale_factors"], float(g["beta"]), float(g["cutoff"]))
    u, _, _ = O.nonbonded(*args)
    np.testing.assert_allclose(u, g["u"], rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(O.nonbonded_energy_dense(*args), g["u"], rtol=1e-12, atol=1e-10)
    u_sub, _, _ = O.nonbonded(*args, atom_idxs=g["atom_idxs"])
    np.testing.assert_allclose(u_sub, g["u_subset"], rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_nonbonded_gradients_match_fd_of_reference_energy(golden, tag):
    g = golden(f"nonbonded_{tag}")
    _, du_dx, du_dp = O.nonbonded(
        g["x"], g["params"], g["box"], g["exclusion_idxs"], g["scale_factors"], float(g["beta"]), float(g["cutoff"])
    )
    scale = max(1.0, np.abs(g["du_dx_fd"]).max())
    np.testing.assert_allclose(du_dx, g["du_dx_fd"], rtol=0, atol=2e-6 * scale)
    # w == cutoff-parked atoms sit on the discontinuity of the cutoff: FD in w is meaningless there
    parked = g["params"][:, 3] >= float(g["cutoff"])
    scale_p = max(1.0, np.abs(g["du_dp_fd"]).max())
    fd = g["du_dp_fd"].copy()
    # LJ is switched off when eps_i == 0 (k_nonbonded.cuh:232; `jnp.where(eps_ij != 0, ...)` nonbonded.py:302), so the
    # reference's derivative w.r.t. that eps is 0 while a finite difference steps across the switch
    fd[g["params"][:, 2] == 0, 2] = 0.0
    np.testing.assert_allclose(du_dp[~parked], fd[~parked], rtol=0, atol=2e-5 * scale_p)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_block_and_pair_list_match_reference_python(golden, tag):
    g = golden(f"nonbonded_{tag}")
    beta, cutoff = float(g["beta"]), float(g["cutoff"])
    u_block, _, _ = O.nonbonded_interaction_group(g["x"], g["params"], g["box"], g["rows"], g["cols"], beta, cutoff)
    np.testing.assert_allclose(u_block, g["u_block"], rtol=1e-12, atol=1e-10)
    u_pairs, _, _ = O.nonbonded_pair_list(g["x"], g["params"], g["box"], g["exclusion_idxs"], g["scale_factors"], beta, cutoff)
    np.testing.assert_allclose(u_pairs, g["u_pairs"], rtol=1e-12, atol=1e-10)


def test_decomposition_all_pairs_minus_exclusions(golden):
    g = golden("nonbonded_b")
    beta, cutoff = float(g["beta"]), float(g["cutoff"])
    N = len(g["x"])
    rows, cols = np.arange(0, 20), np.arange(20, N)
    # all-pairs over everything == all-pairs(rows) + all-pairs(cols) + ixn-group(rows x cols)
    full = O.nonbonded_all_pairs(g["x"], g["params"], g["box"], beta, cutoff)
    a = O.nonbonded_all_pairs(g["x"], g["params"], g["box"], beta, cutoff, atom_idxs=rows)
    b = O.nonbonded_all_pairs(g["x"], g["params"], g["box"], beta, cutoff, atom_idxs=cols)
    c = O.nonbonded_interaction_group(g["x"], g["params"], g["box"], rows, cols, beta, cutoff)
    for k in range(3):
        np.testing.assert_allclose(full[k], a[k] + b[k] + c[k], rtol=1e-10, atol=1e-9)


def test_bonded_match_reference_python(golden):
    g = golden("bonded")
    u, dx, dp = O.harmonic_bond(g["x"], g["bond_params"], g["bond_idxs"])
    np.testing.assert_allclose(u, g["u_bond"], rtol=1e-12)
    np.testing.assert_allclose(dx, g["bond_du_dx_fd"], atol=1e-5 * np.abs(g["bond_du_dx_fd"]).max())
    np.testing.assert_allclose(dp, g["bond_du_dp_fd"], atol=1e-5 * np.abs(g["bond_du_dp_fd"]).max())
    u, dx, dp = O.harmonic_angle(g["x"], g["angle_params"], g["angle_idxs"])
    np.testing.assert_allclose(u, g["u_angle"], rtol=1e-12)
    np.testing.assert_allclose(dx, g["angle_du_dx_fd"], atol=1e-5 * np.abs(g["angle_du_dx_fd"]).max())
    np.testing.assert_allclose(dp, g["angle_du_dp_fd"], atol=1e-5 * np.abs(g["angle_du_dp_fd"]).max())
    u, dx, dp = O.periodic_torsion(g["x"], g["torsion_params"], g["torsion_idxs"])
    np.testing.assert_allclose(u, g["u_torsion"], rtol=1e-12)
    np.testing.assert_allclose(dx, g["torsion_du_dx_fd"], atol=1e-5 * np.abs(g["torsion_du_dx_fd"]).max())
    np.testing.assert_allclose(dp, g["torsion_du_dp_fd"], atol=1e-5 * np.abs(g["torsion_du_dp_fd"]).max())


def test_integrator_matches_reference_python(golden):
    g = golden("integrator")
    ca, cb, cc = O.langevin_coefficients(float(g["temperature"]), float(g["dt"]), float(g["friction"]), g["masses"], boltz=float(g["boltz"]))
    np.testing.assert_allclose(ca, g["ca"], rtol=1e-15)
    np.testing.assert_allclose(cb, g["cb"], rtol=1e-15)
    np.testing.assert_allclose(cc, g["cc"], rtol=1e-15)
    new_x, new_v = O.baoab_step(g["x"], g["v"], g["force"], ca, cb, cc, float(g["dt"]), g["noise"])
    np.testing.assert_allclose(new_x, g["new_x"], rtol=1e-14)
    np.testing.assert_allclose(new_v, g["new_v"], rtol=1e-14)


def test_mixed_precision_step_close_to_f64_step(golden):
    g = golden("integrator")
    T, dt, fr = float(g["temperature"]), float(g["dt"]), float(g["friction"])
    fixed = O.float_to_fixed(-g["force"])  # du_dx = -force
    x32, v32 = O.baoab_step_mixed(g["x"], g["v"], fixed, g["masses"], T, dt, fr, g["noise"].astype(np.float32))
    ca, cb, cc = O.langevin_coefficients(T, dt, fr, g["masses"])
    x64, v64 = O.baoab_step(g["x"], g["v"], g["force"], ca, cb, cc, dt, g["noise"])
    np.testing.assert_allclose(x32, x64, rtol=0, atol=1e-7)
    np.testing.assert_allclose(v32, v64, rtol=2e-6, atol=1e-6)


def test_fixed_point_round_trip():
    v = np.array([0.0, 1.0, -1.0, 1e-11, -1e-11, 123.456, -2.5e6])
    f = O.float_to_fixed(v)
    np.testing.assert_allclose(O.fixed_to_float(f), v, atol=2**-37)
    assert O.float_to_fixed(np.array([0.5 / 2**36]))[0] == 0  # round half to even
    assert O.float_to_fixed(np.array([1.5 / 2**36]))[0] == 2


def test_block_bounds_and_ixn_list_shapes(rng):
    x = rng.uniform(0, 3.0, (100, 3))
    box = np.eye(3) * 3.0
    ctr, ext = O.reference_block_bounds(x, box)
    assert ctr.shape == (4, 3) and (ext >= 0).all()
    ixn, margin = O.reference_ixn_list(x, box, 1.0)
    assert len(ixn) == 4 and margin > 0
    # every atom is its own neighbour in its own row block (d = 0 < cutoff)
    for b, lst in enumerate(ixn):
        for i in range(b * 32, min((b + 1) * 32, 100)):
            assert i in lst
