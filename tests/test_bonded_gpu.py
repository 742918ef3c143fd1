"""GPU: bonded terms vs the oracle (reference tests/test_bonded.py): all output-flag combinations, bitwise repeatability,
the b0 == 0 bond branch, eps-regularised angles, i<->k reversal symmetry, argument validation messages."""

import itertools

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import assert_forces_close, require_reference_ops, round_to_f32

pytestmark = pytest.mark.gpu


def pots():
    from timemachine_b200 import potentials

    return potentials


def ops():
    from timemachine_b200 import custom_ops

    return custom_ops


def check(impl, oracle_fn, x, params, box, precision, rtol32=2e-5):
    rtol = 1e-9 if precision == np.float64 else rtol32
    x, params = round_to_f32(x), round_to_f32(params)
    ref_u, ref_dx, ref_dp = oracle_fn(x, params)
    for want_dx, want_dp, want_u in itertools.product([False, True], repeat=3):
        r1 = impl.execute(x, params, box, want_dx, want_dp, want_u)
        r2 = impl.execute(x, params, box, want_dx, want_dp, want_u)
        for a, b in zip(r1, r2):
            if a is not None:
                np.testing.assert_array_equal(a, b)
        dx, dp, u = r1
        if want_u:
            np.testing.assert_allclose(u, ref_u, rtol=rtol * 10, atol=1e-6)
        if want_dx:
            assert_forces_close(ref_dx, dx, rtol * 10)
        if want_dp:
            assert dp.shape == params.shape
            np.testing.assert_allclose(dp, ref_dp, rtol=rtol * 100, atol=rtol * 100 * max(1.0, np.abs(ref_dp).max()))


def chain(rng, n):
    x = np.cumsum(rng.normal(0, 0.08, (n, 3)), axis=0) + rng.normal(0, 0.02, (n, 3))
    return x, np.eye(3) * 100.0


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_harmonic_bond(precision, rng):
    n = 400
    x, box = chain(rng, n)
    idxs = np.array([(i, i + 1) for i in range(n - 1)] + [(i, i + 5) for i in range(0, n - 5, 3)], dtype=np.int32)
    params = np.stack([rng.uniform(100, 5000, len(idxs)), rng.uniform(0.05, 0.3, len(idxs))], 1)
    params[::5, 1] = 0.0  # b0 == 0 branch
    impl = pots().HarmonicBond(idxs).to_gpu(precision).unbound_impl
    check(impl, lambda x_, p_: O.harmonic_bond(x_, p_, idxs), x, params, box, precision)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_harmonic_angle(precision, rng):
    n = 300
    x, box = chain(rng, n)
    idxs = np.array([(i, i + 1, i + 2) for i in range(n - 2)], dtype=np.int32)
    params = np.stack([rng.uniform(50, 800, len(idxs)), rng.uniform(0.8, 3.0, len(idxs)), rng.choice([0.0, 1e-3, 0.1], len(idxs))], 1)
    impl = pots().HarmonicAngle(idxs).to_gpu(precision).unbound_impl
    check(impl, lambda x_, p_: O.harmonic_angle(x_, p_, idxs), x, params, box, precision)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_harmonic_angle_reversal_is_bitwise_symmetric(precision, rng):
    """tests/test_bonded.py:305-322: listing an angle as (k, j, i) instead of (i, j, k) changes no bit."""
    n = 90
    x, box = chain(rng, n)
    idxs = np.array([(i, i + 1, i + 2) for i in range(n - 2)], dtype=np.int32)
    params = np.stack([rng.uniform(50, 800, len(idxs)), rng.uniform(0.8, 3.0, len(idxs)), np.full(len(idxs), 0.01)], 1)
    a = pots().HarmonicAngle(idxs).to_gpu(precision).unbound_impl.execute(x, params, box)
    b = pots().HarmonicAngle(idxs[:, ::-1].copy()).to_gpu(precision).unbound_impl.execute(x, params, box)
    for u, v in zip(a, b):
        np.testing.assert_array_equal(u, v)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_periodic_torsion(precision, rng):
    n = 300
    x, box = chain(rng, n)
    idxs = np.array([(i, i + 1, i + 2, i + 3) for i in range(n - 3)], dtype=np.int32)
    params = np.stack(
        [rng.uniform(0.5, 30, len(idxs)), rng.uniform(-np.pi, np.pi, len(idxs)), rng.integers(1, 7, len(idxs)).astype(float)], 1
    )
    impl = pots().PeriodicTorsion(idxs).to_gpu(precision).unbound_impl
    check(impl, lambda x_, p_: O.periodic_torsion(x_, p_, idxs), x, params, box, precision, rtol32=1e-4)


def test_empty_bonded_terms():
    box = np.eye(3) * 10
    x = np.zeros((4, 3))
    for cls, width in ((pots().HarmonicBond, 2), (pots().HarmonicAngle, 3), (pots().PeriodicTorsion, 4)):
        impl = cls(np.zeros((0, width), dtype=np.int32)).to_gpu(np.float32).unbound_impl
        dx, dp, u = impl.execute(x, np.zeros((0, 3 if width > 2 else 2)), box)
        assert u == 0.0 and not dx.any() and dp.shape[0] == 0


def test_bonded_validation_messages():
    o = ops()
    with pytest.raises(RuntimeError, match="bond_idxs.size\\(\\) must be exactly 2\\*k!"):
        o.HarmonicBond_f32(np.array([0, 1, 2], dtype=np.int32))
    with pytest.raises(RuntimeError, match="src == dst"):
        o.HarmonicBond_f32(np.array([[1, 1]], dtype=np.int32))
    with pytest.raises(RuntimeError, match="angle_idxs.size\\(\\) must be exactly 3\\*A"):
        o.HarmonicAngle_f32(np.array([0, 1], dtype=np.int32))
    with pytest.raises(RuntimeError, match="angle triplets must be unique"):
        o.HarmonicAngle_f32(np.array([[0, 1, 0]], dtype=np.int32))
    with pytest.raises(RuntimeError, match="torsion_idxs.size\\(\\) must be exactly 4\\*k"):
        o.PeriodicTorsion_f32(np.array([0, 1, 2], dtype=np.int32))
    with pytest.raises(RuntimeError, match="torsion quads must be unique"):
        o.PeriodicTorsion_f32(np.array([[0, 1, 2, 0]], dtype=np.int32))
    bond = o.HarmonicBond_f32(np.array([[0, 1]], dtype=np.int32))
    with pytest.raises(RuntimeError) as e:
        bond.execute(np.zeros((2, 3)), np.zeros((2, 2)), np.eye(3))
    assert str(e.value) == "HarmonicBond::execute_device(): expected P == 2*B, got P=4, 2*B=2"
    angle = o.HarmonicAngle_f32(np.array([[0, 1, 2]], dtype=np.int32))
    with pytest.raises(RuntimeError) as e:
        angle.execute(np.zeros((3, 3)), np.zeros((2, 3)), np.eye(3))
    assert str(e.value) == "HarmonicAngle::execute_device(): expected P == 3*A_, got P=6, 3*A_=3"
    tors = o.PeriodicTorsion_f32(np.array([[0, 1, 2, 3]], dtype=np.int32))
    with pytest.raises(RuntimeError) as e:
        tors.execute(np.zeros((4, 3)), np.zeros((2, 3)), np.eye(3))
    assert str(e.value) == "PeriodicTorsion::execute_device(): expected P == 3*T_, got P=6, 3*T_=3"


@pytest.mark.parametrize("precision,rtol", [(np.float32, 1e-5), (np.float64, 1e-10)])
def test_bonded_against_reference_custom_ops(precision, rtol, rng):
    ref = require_reference_ops()
    suffix = "f32" if precision == np.float32 else "f64"
    n = 200
    x, box = chain(rng, n)
    x = round_to_f32(x)
    cases = [
        ("HarmonicBond", np.array([(i, i + 1) for i in range(n - 1)], dtype=np.int32),
         np.stack([rng.uniform(100, 5000, n - 1), rng.uniform(0.05, 0.3, n - 1)], 1)),
        ("HarmonicAngle", np.array([(i, i + 1, i + 2) for i in range(n - 2)], dtype=np.int32),
         np.stack([rng.uniform(50, 800, n - 2), rng.uniform(0.8, 3.0, n - 2), np.full(n - 2, 0.01)], 1)),
        ("PeriodicTorsion", np.array([(i, i + 1, i + 2, i + 3) for i in range(n - 3)], dtype=np.int32),
         np.stack([rng.uniform(0.5, 30, n - 3), rng.uniform(-3, 3, n - 3), rng.integers(1, 7, n - 3).astype(float)], 1)),
    ]
    for name, idxs, params in cases:
        params = round_to_f32(params)
        rdx, rdp, ru = getattr(ref, f"{name}_{suffix}")(idxs).execute(x, params, box)
        dx, dp, u = getattr(ops(), f"{name}_{suffix}")(idxs).execute(x, params, box)
        assert_forces_close(rdx, dx, rtol, what=name)
        np.testing.assert_allclose(dp, rdp, rtol=rtol * 10, atol=rtol * 10 * max(1.0, np.abs(rdp).max()))
        np.testing.assert_allclose(u, ru, rtol=rtol * 10)
        if precision == np.float32:
            # the f32 kernels reproduce the reference's operation sequence (its FMA contraction is written out in k_bonded.cu)
            np.testing.assert_array_equal(dx, rdx, err_msg=name)
            np.testing.assert_array_equal(dp, rdp, err_msg=name)
            assert u == ru, name


@pytest.mark.parametrize("precision", [np.float32, np.float64])
def test_bonded_far_from_the_origin(precision, rng):
    """Coordinates are never wrapped into the box, atoms drift nanometres away from the origin.  The reference forms
    x_i - x_j in double and rounds the DIFFERENCE to the kernel's type (k_harmonic_bond.cuh:26, k_harmonic_angle.cuh:43-44,
    k_periodic_torsion.cuh:48-50): the f32 kernels keep ~1e-8 nm on a bond vector wherever the molecule sits.  Rounding the
    coordinates first would lose ~1e-6 nm at 10 nm, a 1e-3 relative error on a stiff bond.  Checked against the f64 oracle
    and, on f64-only-representable coordinates, bitwise against the compiled reference."""
    n = 200
    x, box = chain(rng, n)
    x = x + np.array([11.3, -9.7, 10.1])  # NOT rounded to f32: the low bits of the doubles matter here
    cases = [
        ("HarmonicBond", O.harmonic_bond, np.array([(i, i + 1) for i in range(n - 1)], dtype=np.int32),
         np.stack([np.full(n - 1, 462750.4), np.linalg.norm(x[1:] - x[:-1], axis=1) + rng.normal(0, 0.003, n - 1)], 1)),
        ("HarmonicAngle", O.harmonic_angle, np.array([(i, i + 1, i + 2) for i in range(n - 2)], dtype=np.int32),
         np.stack([rng.uniform(50, 800, n - 2), rng.uniform(0.8, 3.0, n - 2), np.full(n - 2, 0.01)], 1)),
        ("PeriodicTorsion", O.periodic_torsion, np.array([(i, i + 1, i + 2, i + 3) for i in range(n - 3)], dtype=np.int32),
         np.stack([rng.uniform(0.5, 30, n - 3), rng.uniform(-3, 3, n - 3), rng.integers(1, 7, n - 3).astype(float)], 1)),
    ]
    suffix = "f32" if precision == np.float32 else "f64"
    from tests.common import load_reference_ops

    ref = load_reference_ops()
    for name, oracle_fn, idxs, params in cases:
        params = round_to_f32(params)
        dx, dp, u = getattr(ops(), f"{name}_{suffix}")(idxs).execute(x, params, box)
        _, odx, _ = oracle_fn(x, params, idxs)
        # stiff bonds: |r - b0| ~ 3e-3 nm, so a 1e-6 nm error in the bond vector would show up as 3e-4 here
        # f32 arithmetic on an exact bond vector: ~3e-5 of the force on these stiff bonds (sqrt and the r - b0 difference)
        assert_forces_close(odx, dx, 1e-4 if precision == np.float32 else 1e-9, what=name + " vs f64 oracle, offset 10 nm")
        if ref is not None:
            rdx, rdp, ru = getattr(ref, f"{name}_{suffix}")(idxs).execute(x, params, box)
            if precision == np.float32:
                # the f32 kernels reproduce the reference's operation sequence
                np.testing.assert_array_equal(dx, rdx, err_msg=name + ": forces are not bitwise the compiled reference's")
            assert_forces_close(rdx, dx, 1e-10, what=name + " vs compiled reference, offset 10 nm")
            np.testing.assert_allclose(u, ru, rtol=1e-4 if precision == np.float32 else 1e-9)
