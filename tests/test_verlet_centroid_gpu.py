"""GPU: the two classes of custom_ops.pyi outside SURVEY.md §8's tables that complete the module - CentroidRestraint_{f32,f64}
and VelocityVerletIntegrator - modelled on the reference's tests (tests/test_bonded.py:22-95,
tests/test_velocity_verlet_integrator.py, tests/test_md.py:250-300): against the CPU oracle (pinned to the reference's Python,
tests/test_oracle_verlet_centroid.py), bitwise against the compiled reference, reversibility, the initialize / finalize
protocol and its messages."""

from pathlib import Path

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import assert_forces_close, require_reference_ops, round_to_f32, water_box

pytestmark = pytest.mark.gpu
G = dict(np.load(Path(__file__).parent / "golden" / "verlet_centroid.npz"))
BOX = np.eye(3) * 3.0
NO_PARAMS = np.zeros(0)


def ops():
    from timemachine_b200 import custom_ops

    return custom_ops


@pytest.mark.parametrize("suffix,rtol", [("f64", 1e-10), ("f32", 1e-5)])
@pytest.mark.parametrize("b0", [None, 0.0])
def test_centroid_restraint_against_oracle(suffix, rtol, b0):
    b0 = float(G["b0"]) if b0 is None else b0
    kb = float(G["kb"])
    impl = getattr(ops(), f"CentroidRestraint_{suffix}")(G["group_a"], G["group_b"], kb, b0)
    ref_u, ref_dx = O.centroid_restraint(G["x"], G["group_a"], G["group_b"], kb, b0)
    dx, dp, u = impl.execute(G["x"], NO_PARAMS, BOX)
    np.testing.assert_allclose(u, ref_u, rtol=rtol)
    assert_forces_close(ref_dx, dx, rtol)
    assert dp.shape == (0,)
    # flag combinations; the energy is overwritten, the forces are repeatable bit for bit
    dx2, _, u2 = impl.execute(G["x"], NO_PARAMS, BOX, True, False, True)
    np.testing.assert_array_equal(dx, dx2)
    assert u == u2
    assert impl.execute(G["x"], NO_PARAMS, BOX, False, False, True)[2] == u
    np.testing.assert_array_equal(impl.execute(G["x"], NO_PARAMS, BOX, True, False, False)[0], dx)


@pytest.mark.parametrize("suffix", ["f32", "f64"])
@pytest.mark.parametrize("b0", [0.35, 0.0])
def test_centroid_restraint_bitwise_against_reference(suffix, b0):
    ref = require_reference_ops()
    rng = np.random.default_rng(11)
    for n_a, n_b in [(1, 1), (9, 14), (300, 700)]:  # the last one needs more than one pass of the CTA over the atoms
        n = n_a + n_b + 5
        x = rng.uniform(-1, 4, (n, 3))
        perm = rng.permutation(n)
        ga, gb = perm[:n_a].astype(np.int32), perm[n_a : n_a + n_b].astype(np.int32)
        mine = getattr(ops(), f"CentroidRestraint_{suffix}")(ga, gb, 77.7, b0)
        theirs = getattr(ref, f"CentroidRestraint_{suffix}")(ga, gb, 77.7, b0)
        dx, _, u = mine.execute(x, NO_PARAMS, BOX)
        rdx, _, ru = theirs.execute(x, NO_PARAMS, BOX)
        np.testing.assert_array_equal(dx, rdx)
        assert u == ru


def test_centroid_restraint_through_the_dataclass_and_in_a_sum():
    from timemachine_b200 import potentials as P

    pot = P.CentroidRestraint(G["group_a"], G["group_b"], float(G["kb"]), float(G["b0"]))
    u = pot.to_gpu(np.float64)(G["x"], NO_PARAMS, BOX)
    assert u == pytest.approx(float(G["u_b0"]), rel=1e-10)
    bond_idxs = np.array([[0, 1], [2, 3]], dtype=np.int32)
    bond_params = np.array([[1000.0, 0.2], [500.0, 0.3]])
    summed = ops().SummedPotential(
        [ops().HarmonicBond_f64(bond_idxs), pot.to_gpu(np.float64).unbound_impl], [bond_params.size, 0], False
    )
    dx, dp, us = summed.execute(G["x"], bond_params.reshape(-1), BOX)
    dxb, dpb, ub = ops().HarmonicBond_f64(bond_idxs).execute(G["x"], bond_params, BOX)
    dxc, _, uc = pot.to_gpu(np.float64).unbound_impl.execute(G["x"], NO_PARAMS, BOX)
    np.testing.assert_array_equal(dx, dxb + dxc)  # fixed-point sums: exact whatever the order
    np.testing.assert_array_equal(dp, dpb.reshape(-1))
    assert us == pytest.approx(ub + uc, rel=1e-14)


# ---- velocity Verlet -------------------------------------------------------------------------------------------------
BETA, CUTOFF = 2.0, 1.2


def water_potentials(n_waters, seed, suffix, module=None):
    """Flexible waters (tests/common.py water_box): bonds, angles, all-pairs nonbonded and its exclusions, one bound
    potential each, built from the raw custom_ops classes so that the compiled reference module can be driven identically."""
    o = module if module is not None else ops()
    s = water_box(n_waters, seed=seed)
    x0 = round_to_f32(s["x"])
    nb_params = round_to_f32(s["params"])
    bps = [
        o.BoundPotential(getattr(o, f"HarmonicBond_{suffix}")(s["bond_idxs"]), round_to_f32(s["bond_params"])),
        o.BoundPotential(getattr(o, f"HarmonicAngle_{suffix}")(s["angle_idxs"]), round_to_f32(s["angle_params"])),
        o.BoundPotential(getattr(o, f"NonbondedAllPairs_{suffix}")(s["N"], BETA, CUTOFF), nb_params),
        o.BoundPotential(
            getattr(o, f"NonbondedExclusions_{suffix}")(s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF), nb_params
        ),
    ]
    return bps, x0, s["box"], s["masses"]


def small_water_system(suffix="f32", n_waters=700, seed=3):
    return water_potentials(n_waters, seed, suffix)


def build_context(bps, x0, v0, box, dt, masses):
    from timemachine_b200 import lib

    intg = lib.VelocityVerletIntegrator(dt, masses).impl()
    return ops().Context(x0, v0, box, intg, bps), intg


def test_velocity_verlet_matches_the_oracle_and_the_reference_protocol():
    """tests/test_velocity_verlet_integrator.py:106-141: n steps through Context.multiple_steps are the reference's Python
    integrator run for n + 1 (the context brackets them with the two half steps); here against the f64 restatement of the
    compiled scheme with forces from the same potentials."""
    bps, x0, box, masses = small_water_system()
    rng = np.random.default_rng(2022)
    v0 = rng.normal(0, 0.3, x0.shape)
    dt, n = 1.0e-3, 20
    ctx, _ = build_context(bps, x0, v0, box, dt, masses)

    def du_dx(x):
        return sum(bp.execute(x, box, True, False)[0] for bp in bps)

    frames, x_end, v_end = O.velocity_verlet_f64(du_dx, x0, v0, -dt / masses, dt, n)
    xs, boxes = ctx.multiple_steps(n, 1)
    assert xs.shape == (n, len(x0), 3)
    np.testing.assert_allclose(xs, frames, atol=1e-9)
    np.testing.assert_allclose(ctx.get_x_t(), x_end, atol=1e-9)
    np.testing.assert_allclose(ctx.get_v_t(), v_end, atol=1e-7)
    # and against the Python class's fixed-point form (integrator.py:169-201), as the reference's own test does
    ref_xs, ref_vs = O.velocity_verlet_multiple_steps(lambda x: -du_dx(x), x0, v0, masses, dt, n + 1)
    np.testing.assert_allclose(ref_xs[1:-1], xs, atol=1e-5)
    np.testing.assert_allclose(ref_vs[-1], ctx.get_v_t(), atol=1e-5)


@pytest.mark.parametrize("n_steps", [1, 10, 100])
def test_velocity_verlet_is_reversible(n_steps):
    """tests/test_velocity_verlet_integrator.py:21-105: integrate, flip v, integrate, flip v returns to the start - through
    initialize / step / finalize and through multiple_steps."""
    # f64 potentials: an f32 kernel rounds the coordinates it reads, and on the way back a coordinate that differs in its
    # 13th digit occasionally rounds to the neighbouring float, which shows up as 1e-6 in a velocity (seen at 100 steps on
    # 2100 atoms); the reference's test gets away with f32 on a 60-atom ligand in vacuum
    bps, x0, box, masses = small_water_system("f64")
    rng = np.random.default_rng(2022)
    v0 = rng.normal(0, 0.3, x0.shape)
    ctx, _ = build_context(bps, x0, v0, box, 1.0e-3, masses)

    def by_steps(x, v):
        ctx.set_x_t(x)
        ctx.set_v_t(v)
        ctx.initialize()
        for _ in range(n_steps):
            ctx.step()
        ctx.finalize()
        return ctx.get_x_t(), ctx.get_v_t()

    def by_multiple_steps(x, v):
        ctx.set_x_t(x)
        ctx.set_v_t(v)
        xs, _ = ctx.multiple_steps(n_steps)
        return xs[-1], ctx.get_v_t()

    for update in (by_steps, by_multiple_steps):
        x1, v1 = update(x0, v0)
        x2, v2 = update(x1, -v1)
        np.testing.assert_allclose(x2, x0, atol=1e-10)
        np.testing.assert_allclose(-v2, v0, atol=1e-10)
        assert not np.allclose(x1, x0, atol=1e-10) and not np.allclose(v1, v0, atol=1e-10)
    # the two drivers are the same arithmetic
    xa, va = by_steps(x0, v0)
    xb, vb = by_multiple_steps(x0, v0)
    np.testing.assert_array_equal(xa, xb)
    np.testing.assert_array_equal(va, vb)


def test_velocity_verlet_initialize_finalize_protocol():
    bps, x0, box, masses = small_water_system()
    ctx, _ = build_context(bps, x0, np.zeros_like(x0), box, 1.0e-3, masses)
    with pytest.raises(RuntimeError, match="not initialized"):
        ctx.finalize()
    ctx.initialize()
    with pytest.raises(RuntimeError, match="initialized twice"):
        ctx.initialize()
    ctx.finalize()
    with pytest.raises(RuntimeError, match="not initialized"):
        ctx.finalize()
    # local MD draws its temperature from a Langevin integrator unless it was configured explicitly (test_md.py:279-283)
    with pytest.raises(RuntimeError, match="integrator must be LangevinIntegrator."):
        ctx.multiple_steps_local(10, np.array([0], dtype=np.int32))
    # a Langevin context accepts initialize / finalize as no-ops
    lang = ops().LangevinIntegrator(masses, 300.0, 1.0e-3, 1.0, 1)
    lctx = ops().Context(x0, np.zeros_like(x0), box, lang, bps)
    lctx.finalize()
    lctx.initialize()
    lctx.initialize()


def test_velocity_verlet_conserves_energy():
    bps, x0, box, masses = small_water_system("f64")
    rng = np.random.default_rng(4)
    v0 = rng.normal(0, 1, x0.shape) * np.sqrt(0.008314462618 * 300.0 / masses)[:, None]
    ctx, _ = build_context(bps, x0, v0, box, 0.25e-3, masses)

    def total_energy():
        x, v = ctx.get_x_t(), ctx.get_v_t()
        return sum(bp.execute(x, box, False, True)[1] for bp in bps) + 0.5 * np.sum(masses[:, None] * v * v)

    ctx.multiple_steps(200)  # leave the start-up transient of the synthetic geometry behind
    e0 = total_energy()
    kinetic = 0.5 * np.sum(masses[:, None] * ctx.get_v_t() ** 2)
    drift = []
    for _ in range(5):
        ctx.multiple_steps(200)
        drift.append(total_energy() - e0)
    assert np.max(np.abs(drift)) < 1e-2 * kinetic, (drift, kinetic)


def test_velocity_verlet_bitwise_against_reference():
    """f32 potentials are bitwise the reference's, the integrator's three statements are the reference's FMAs: so is the
    trajectory."""
    ref = require_reference_ops()
    bps, x0, box, masses = water_potentials(700, 3, "f32")
    rbps, _, _, _ = water_potentials(700, 3, "f32", module=ref)
    rng = np.random.default_rng(8)
    v0 = rng.normal(0, 0.3, x0.shape)
    dt = 1.5e-3
    cbs = -dt / masses
    mine = ops().Context(x0, v0, box, ops().VelocityVerletIntegrator(dt, cbs), bps)
    theirs = ref.Context(x0, v0, box, ref.VelocityVerletIntegrator(dt, cbs), rbps)
    xs, _ = mine.multiple_steps(50, 10)
    rxs, _ = theirs.multiple_steps(50, 10)
    np.testing.assert_array_equal(xs, rxs)
    np.testing.assert_array_equal(mine.get_v_t(), theirs.get_v_t())
    mine.initialize()
    theirs.initialize()
    for _ in range(3):
        mine.step()
        theirs.step()
    mine.finalize()
    theirs.finalize()
    np.testing.assert_array_equal(mine.get_x_t(), theirs.get_x_t())
    np.testing.assert_array_equal(mine.get_v_t(), theirs.get_v_t())
