"""GPU: composition and batch API (reference tests/test_potentials.py): BoundPotential lifetime/set_params, Summed /
Fanout potentials, execute_batch / execute_batch_sparse shapes and values."""

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import assert_forces_close, round_to_f32, water_box

pytestmark = pytest.mark.gpu
BETA, CUTOFF = 2.0, 1.2


def pots():
    from timemachine_b200 import potentials

    return potentials


def ops():
    from timemachine_b200 import custom_ops

    return custom_ops


@pytest.fixture(scope="module")
def system():
    s = water_box(150, seed=3)
    s["x"] = round_to_f32(s["x"])
    s["params"] = round_to_f32(s["params"])
    return s


def summed_system(s, precision, parallel=True):
    P = pots()
    nb = P.Nonbonded(s["N"], s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF)
    pot = P.SummedPotential(
        [P.HarmonicBond(s["bond_idxs"]), P.HarmonicAngle(s["angle_idxs"]), nb],
        [s["bond_params"], s["angle_params"], s["params"]],
        parallel,
    )
    flat = np.concatenate([s["bond_params"].reshape(-1), s["angle_params"].reshape(-1), s["params"].reshape(-1)])
    return pot.to_gpu(precision).unbound_impl, flat


def oracle_total(s, x=None):
    x = s["x"] if x is None else x
    ub, dxb, dpb = O.harmonic_bond(x, s["bond_params"], s["bond_idxs"])
    ua, dxa, dpa = O.harmonic_angle(x, s["angle_params"], s["angle_idxs"])
    un, dxn, dpn = O.nonbonded(x, s["params"], s["box"], s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF)
    return ub + ua + un, dxb + dxa + dxn, np.concatenate([dpb.reshape(-1), dpa.reshape(-1), dpn.reshape(-1)])


@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("parallel", [True, False])
def test_summed_potential(system, precision, parallel):
    impl, flat = summed_system(system, precision, parallel)
    ref_u, ref_dx, ref_dp = oracle_total(system)
    dx, dp, u = impl.execute(system["x"], flat, system["box"])
    rtol = 1e-8 if precision == np.float64 else 1e-4
    np.testing.assert_allclose(u, ref_u, rtol=rtol, atol=5e-3 if precision == np.float32 else 1e-7)
    assert_forces_close(ref_dx, dx, rtol)
    assert dp.shape == flat.shape
    assert_forces_close(ref_dp.reshape(-1, 1), dp.reshape(-1, 1), rtol * 100, what="du_dp")
    # streams on/off must not change a bit
    impl2, _ = summed_system(system, precision, not parallel)
    for a, b in zip(impl.execute(system["x"], flat, system["box"]), impl2.execute(system["x"], flat, system["box"])):
        np.testing.assert_array_equal(a, b)
    assert len(impl.get_potentials()) == 3
    with pytest.raises(RuntimeError, match="SummedPotential::execute_device\\(\\): expected"):
        impl.execute(system["x"], flat[:-1], system["box"])


def test_summed_potential_validation():
    o = ops()
    bond = o.HarmonicBond_f32(np.array([[0, 1]], dtype=np.int32))
    with pytest.raises(RuntimeError, match="number of potentials != number of parameter sizes"):
        o.SummedPotential([bond], [2, 3])


def test_bound_potential(system):
    impl, flat = summed_system(system, np.float32)
    o = ops()
    bp = o.BoundPotential(impl, flat)
    assert bp.size() == flat.size and bp.get_potential() is impl
    dx_u = impl.execute(system["x"], flat, system["box"], True, False, True)
    dx_b, u_b = bp.execute(system["x"], system["box"])
    np.testing.assert_array_equal(dx_u[0], dx_b)
    assert dx_u[2] == u_b
    assert bp.execute(system["x"], system["box"], compute_du_dx=False)[0] is None
    assert bp.execute(system["x"], system["box"], compute_u=False)[1] is None
    fixed = bp.execute_fixed(system["x"], system["box"])
    assert fixed.dtype == np.uint64 and np.isclose(np.int64(fixed[0]) / 2**36, u_b)
    # set_params swaps the parameter set in place (how HREX moves a replica between states)
    flat2 = flat.copy()
    flat2[-4 * system["N"] :: 4] *= 0.5  # halve every charge
    bp.set_params(flat2)
    dx2, u2 = bp.execute(system["x"], system["box"])
    ref = impl.execute(system["x"], flat2, system["box"], True, False, True)
    np.testing.assert_array_equal(dx2, ref[0])
    assert u2 == ref[2] and u2 != u_b
    with pytest.raises(RuntimeError, match="parameter size is not equal to device buffer size"):
        bp.set_params(flat[:-1])
    # the bound potential keeps its potential alive
    del impl
    assert np.isfinite(bp.execute(system["x"], system["box"])[1])


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_execute_batch(system, precision, rng):
    impl = pots().NonbondedAllPairs(system["N"], BETA, CUTOFF).to_gpu(precision).unbound_impl
    C, Pb = 3, 2
    coords = np.stack([system["x"] + rng.normal(0, 0.01, system["x"].shape) for _ in range(C)])
    boxes = np.stack([system["box"] * (1 + 0.01 * i) for i in range(C)])
    params = np.stack([system["params"], system["params"] * np.array([0.5, 1.0, 1.0, 1.0])])
    dx, dp, u = impl.execute_batch(coords, params, boxes, True, True, True)
    assert dx.shape == (C, Pb, system["N"], 3) and dp.shape == (C, Pb, system["N"], 4) and u.shape == (C, Pb)
    for i in range(C):
        for j in range(Pb):
            sdx, sdp, su = impl.execute(coords[i], params[j], boxes[i])
            np.testing.assert_array_equal(dx[i, j], sdx)
            np.testing.assert_array_equal(dp[i, j], sdp)
            assert u[i, j] == su
    dx, dp, u = impl.execute_batch(coords, params, boxes, False, False, True)
    assert dx is None and dp is None and u.shape == (C, Pb)
    with pytest.raises(RuntimeError, match="coords and boxes must have 3 dimensions"):
        impl.execute_batch(coords[0], params, boxes, True, True, True)
    with pytest.raises(RuntimeError, match="number of batches of coords and boxes don't match"):
        impl.execute_batch(coords, params, boxes[:2], True, True, True)
    with pytest.raises(RuntimeError, match="parameters must have at least 2 dimensions"):
        impl.execute_batch(coords, params.reshape(-1), boxes, True, True, True)


def test_execute_batch_sparse(system, rng):
    impl = pots().NonbondedAllPairs(system["N"], BETA, CUTOFF).to_gpu(np.float32).unbound_impl
    C = 3
    coords = np.stack([system["x"] + rng.normal(0, 0.01, system["x"].shape) for _ in range(C)])
    boxes = np.stack([system["box"]] * C)
    params = np.stack([system["params"] * np.array([s, 1.0, 1.0, 1.0]) for s in (1.0, 0.7, 0.3)])
    cidx = np.array([0, 0, 1, 2, 2], dtype=np.uint32)
    pidx = np.array([0, 1, 1, 1, 2], dtype=np.uint32)
    dx, dp, u = impl.execute_batch_sparse(coords, params, boxes, cidx, pidx, True, True, True)
    assert dx.shape == (5, system["N"], 3) and dp.shape == (5, system["N"], 4) and u.shape == (5,)
    for k, (i, j) in enumerate(zip(cidx, pidx)):
        sdx, sdp, su = impl.execute(coords[i], params[j], boxes[i])
        np.testing.assert_array_equal(dx[k], sdx)
        np.testing.assert_array_equal(dp[k], sdp)
        assert u[k] == su
    with pytest.raises(RuntimeError, match="coords_batch_idxs contains an index that is out of bounds"):
        impl.execute_batch_sparse(coords, params, boxes, np.array([5], dtype=np.uint32), np.array([0], dtype=np.uint32), True, True, True)
    with pytest.raises(RuntimeError, match="must have the same length"):
        impl.execute_batch_sparse(coords, params, boxes, cidx, pidx[:2], True, True, True)


def test_bound_execute_batch(system, rng):
    impl, flat = summed_system(system, np.float32)
    bp = ops().BoundPotential(impl, flat)
    coords = np.stack([system["x"] + rng.normal(0, 0.01, system["x"].shape) for _ in range(3)])
    boxes = np.stack([system["box"]] * 3)
    dx, u = bp.execute_batch(coords, boxes, True, True)
    assert dx.shape == (3, system["N"], 3) and u.shape == (3,)
    for i in range(3):
        sdx, su = bp.execute(coords[i], boxes[i])
        np.testing.assert_array_equal(dx[i], sdx)
        assert u[i] == su


def test_energy_only_and_force_only_paths_agree(system):
    """The 8 kernel variants must agree with each other bitwise on the outputs they share."""
    impl, flat = summed_system(system, np.float32)
    full = impl.execute(system["x"], flat, system["box"], True, True, True)
    only_u = impl.execute(system["x"], flat, system["box"], False, False, True)
    only_x = impl.execute(system["x"], flat, system["box"], True, False, False)
    only_p = impl.execute(system["x"], flat, system["box"], False, True, False)
    assert full[2] == only_u[2]
    np.testing.assert_array_equal(full[0], only_x[0])
    np.testing.assert_array_equal(full[1], only_p[1])
    np.testing.assert_array_equal(impl.execute_du_dx(system["x"], flat, system["box"]), full[0])
