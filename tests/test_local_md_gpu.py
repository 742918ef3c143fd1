"""GPU: local MD (reference context.cu:90-214, local_md_potentials.cu, tests/test_local_md.py).

  * one local step against its definition: forces of the FULL system's potentials plus the flat-bottom restraints to the
    reference atom, integrated for the free atoms only - bit for bit (fixed-point sums do not care that the local
    potentials split the pairs into free-free and free-frozen); frozen atoms do not move at all;
  * the selected shell against the reference's (same cuRAND / mt19937 seeds) and the trajectory against the compiled
    reference's `multiple_steps_local`;
  * explicit selections, argument validation with the reference's messages, and that a context is unchanged for global MD
    after local MD."""

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import load_reference_ops, round_to_f32, water_box

pytestmark = pytest.mark.gpu
BETA, CUTOFF = 2.0, 1.2
TEMPERATURE, DT = 300.0, 1.5e-3


def mods():
    from timemachine_b200 import custom_ops, lib, potentials

    return custom_ops, lib, potentials


@pytest.fixture(scope="module")
def system():
    s = water_box(700, seed=17)  # 2100 atoms, box 2.76 nm
    s["x"] = round_to_f32(s["x"])
    s["params"] = round_to_f32(s["params"])
    rng = np.random.default_rng(3)
    s["v"] = rng.normal(0, 1.0, (s["N"], 3)) * np.sqrt(0.008314462618 * TEMPERATURE / s["masses"])[:, None]
    return s


def make_context(s, friction=0.0, seed=2024, module=None):
    """The water box as [HarmonicBond, HarmonicAngle, Nonbonded] bound potentials (separately bound, as the reference's
    tests/test_local_md.py does) on this library or, with `module`, on the compiled reference."""
    ops, lib, P = mods()
    N = s["N"]
    if module is None:
        bps = [
            P.HarmonicBond(s["bond_idxs"]).bind(s["bond_params"]).to_gpu(np.float32).bound_impl,
            P.HarmonicAngle(s["angle_idxs"]).bind(s["angle_params"]).to_gpu(np.float32).bound_impl,
            P.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF).bind(s["params"]).to_gpu(np.float32).bound_impl,
        ]
        intg = ops.LangevinIntegrator(s["masses"], TEMPERATURE, DT, friction, seed)
        return ops.Context(s["x"], s["v"], s["box"], intg, bps)
    ref = module
    nb = ref.FanoutSummedPotential(
        [ref.NonbondedAllPairs_f32(N, BETA, CUTOFF, None, False, 0.1), ref.NonbondedExclusions_f32(s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF)], True
    )
    bps = [
        ref.BoundPotential(ref.HarmonicBond_f32(s["bond_idxs"]), s["bond_params"]),
        ref.BoundPotential(ref.HarmonicAngle_f32(s["angle_idxs"]), s["angle_params"]),
        ref.BoundPotential(nb, s["params"]),
    ]
    intg = ref.LangevinIntegrator(s["masses"], TEMPERATURE, DT, friction, seed)
    return ref.Context(s["x"], s["v"], s["box"], intg, bps)


def distances_to(s, x, ref_idx):
    d = x - x[ref_idx]
    L = np.diag(s["box"])
    d -= L * np.rint(d / L)
    return np.linalg.norm(d, axis=1)


def test_one_local_step_is_the_full_forces_plus_restraint_on_the_free_atoms(system):
    ops, lib, P = mods()
    s = system
    N, ref_idx, radius, k = s["N"], 300, 0.7, 2000.0
    ctx = make_context(s)
    xs, boxes = ctx.multiple_steps_local(1, np.array([ref_idx], dtype=np.int32), radius=radius, k=k, seed=11)
    assert xs.shape == (1, N, 3) and boxes.shape == (1, 3, 3)
    free = ctx.local_md_free_idxs()
    frozen = np.setdiff1d(np.arange(N), free)
    assert ref_idx in frozen and 20 < len(free) < N - 1
    # the shell: everything inside the radius is free, nothing far outside is (k = 2000: exp(-U/kT) < 1e-9 beyond +0.4 nm)
    d = distances_to(s, s["x"], ref_idx)
    inside = np.flatnonzero((d < radius) & (np.arange(N) != ref_idx))
    assert np.isin(inside, free).all() and (d[free] < radius + 0.4).all()
    # frozen atoms: bitwise where they were, velocities untouched
    np.testing.assert_array_equal(xs[0][frozen], s["x"][frozen])
    np.testing.assert_array_equal(ctx.get_v_t()[frozen], s["v"][frozen])

    # definition: full-system forces (fixed point) + flat-bottom restraints free_i - reference
    fixed = np.zeros((N, 3), dtype=np.int64)
    full = make_context(s)
    for bp in full.get_potentials():
        fixed += np.rint(bp.execute(s["x"], s["box"])[0] * 2.0**36).astype(np.int64)
    bonds = np.stack([np.full(len(free), ref_idx), free], 1).astype(np.int32)
    fb_params = np.tile([k, 0.0, radius], (len(free), 1))
    fb = P.FlatBottomBond(bonds).to_gpu(np.float32).unbound_impl
    fixed += np.rint(fb.execute(s["x"], fb_params, s["box"])[0] * 2.0**36).astype(np.int64)
    x1, v1 = O.baoab_step_mixed(s["x"], s["v"], fixed.view(np.uint64), s["masses"], TEMPERATURE, DT, 0.0, np.zeros((N, 3), np.float32))
    np.testing.assert_array_equal(xs[0][free], x1[free])
    np.testing.assert_array_equal(ctx.get_v_t()[free], v1[free])


def test_shell_and_trajectory_match_the_compiled_reference(system):
    ref = load_reference_ops()
    if ref is None:
        pytest.skip("oracle/_ref/custom_ops*.so not built")
    s = system
    N = s["N"]
    local_idxs = np.array([10, 900, 1500], dtype=np.int32)
    for seed in (5, 6):
        ours = make_context(s)
        theirs = make_context(s, module=ref)
        xs, _ = ours.multiple_steps_local(20, local_idxs, store_x_interval=10, radius=0.6, k=5000.0, seed=seed)
        rxs, _ = theirs.multiple_steps_local(20, local_idxs, 10, 0.6, 5000.0, seed)
        assert xs.shape == rxs.shape == (2, N, 3)
        moved = np.flatnonzero(np.any(xs[-1] != s["x"], axis=1))
        rmoved = np.flatnonzero(np.any(rxs[-1] != s["x"], axis=1))
        np.testing.assert_array_equal(moved, rmoved)  # the same reference atom and the same shell were drawn
        np.testing.assert_array_equal(moved, ours.local_md_free_idxs())
        # 20 steps at friction 0: deterministic, and every f32 kernel involved (bonds, angles, nonbonded tiles and pair lists,
        # the flat-bottom restraint) follows the reference's rounded-operation sequence: the same bits
        np.testing.assert_array_equal(xs, rxs)
        np.testing.assert_array_equal(ours.get_v_t(), theirs.get_v_t())


def test_local_selection_and_context_is_unchanged_afterwards(system):
    s = system
    N = s["N"]
    ctx = make_context(s, friction=1.0)
    ref_idx = 42
    d = distances_to(s, s["x"], ref_idx)
    selection = np.flatnonzero((d < 0.5) & (np.arange(N) != ref_idx)).astype(np.int32)
    xs, _ = ctx.multiple_steps_local_selection(15, ref_idx, selection, store_x_interval=5, radius=0.5, k=1000.0)
    assert xs.shape == (3, N, 3)
    moved = np.flatnonzero(np.any(xs[-1] != s["x"], axis=1))
    np.testing.assert_array_equal(moved, selection)
    np.testing.assert_array_equal(ctx.local_md_free_idxs(), selection)
    # the restraint holds the shell together: nobody ends up far outside the radius
    assert (distances_to(s, xs[-1], ref_idx)[selection] < 0.5 + 0.15).all()

    # global MD after local MD == global MD on a fresh context (the all-pairs potential got its atoms back)
    ctx.set_x_t(s["x"])
    ctx.set_v_t(s["v"])
    ctx.get_integrator().set_step(0)
    a, _ = ctx.multiple_steps(30)
    fresh = make_context(s, friction=1.0)
    b, _ = fresh.multiple_steps(30)
    np.testing.assert_array_equal(a, b)
    # and local MD can be used again, with the default arguments of the reference API
    xs2, _ = ctx.multiple_steps_local(5, np.array([ref_idx], dtype=np.int32))
    assert xs2.shape == (1, N, 3) and np.isfinite(xs2).all()


def test_local_md_validation(system):
    s = system
    ctx = make_context(s)
    idxs = np.array([1], dtype=np.int32)
    with pytest.raises(RuntimeError, match="local steps must be at least one"):
        ctx.multiple_steps_local(0, idxs)
    with pytest.raises(RuntimeError, match="store_x_interval must be greater than or equal to zero"):
        ctx.multiple_steps_local(10, idxs, store_x_interval=-1)
    with pytest.raises(RuntimeError, match="radius must be greater or equal to 0.100000"):
        ctx.multiple_steps_local(10, idxs, radius=0.01)
    with pytest.raises(RuntimeError, match="k must be at least one"):
        ctx.multiple_steps_local(10, idxs, k=0.5)
    with pytest.raises(RuntimeError, match="k must be less than than 1e\\+06"):
        ctx.multiple_steps_local(10, idxs, k=1e7)
    with pytest.raises(RuntimeError, match="indices can't be empty"):
        ctx.multiple_steps_local(10, np.array([], dtype=np.int32))
    with pytest.raises(RuntimeError, match="atom indices must be unique"):
        ctx.multiple_steps_local(10, np.array([1, 1], dtype=np.int32))
    with pytest.raises(RuntimeError, match=f"index values must be less than N\\({s['N']}\\)"):
        ctx.multiple_steps_local(10, np.array([s["N"]], dtype=np.int32))
    with pytest.raises(RuntimeError, match="reference idx must be at least 0 and less than"):
        ctx.multiple_steps_local_selection(10, -1, idxs)
    with pytest.raises(RuntimeError, match="reference idx must not be in selection idxs"):
        ctx.multiple_steps_local_selection(10, 1, idxs)
    # three atoms far from each other, a small radius and a stiff restraint: nobody is selected
    ops, lib, P = mods()
    x3 = np.array([[1.0, 1.0, 1.0], [3.0, 3.0, 3.0], [5.0, 1.0, 4.0]])
    p3 = np.array([[0.5, 0.15, 0.3, 0.0]] * 3)
    nb3 = P.NonbondedAllPairs(3, BETA, CUTOFF).bind(p3).to_gpu(np.float32).bound_impl
    lonely = ops.Context(x3, np.zeros_like(x3), np.eye(3) * 6.0, ops.LangevinIntegrator(np.ones(3) * 12.0, TEMPERATURE, DT, 0.0, 1), [nb3])
    with pytest.raises(RuntimeError, match="LocalMDPotentials setup has no free particles selected"):
        lonely.multiple_steps_local(5, np.array([0], dtype=np.int32), radius=0.1, k=1e6, seed=1)
    # ... and the potential is usable afterwards
    assert np.isfinite(lonely.multiple_steps(3)[0]).all()
    ctx.setup_local_md(TEMPERATURE, True)
    ctx.setup_local_md(TEMPERATURE, True)  # same parameters: fine
    with pytest.raises(RuntimeError, match="local md configured with different parameters"):
        ctx.setup_local_md(TEMPERATURE + 1.0, True)


def test_unfrozen_reference_variant(system):
    """setup_local_md(temperature, freeze_reference=False): the reference atom moves too, the frozen shell is tied to it by
    a LogFlatBottomBond.  One step against the definition (bitwise), then against the compiled reference."""
    ops, lib, P = mods()
    s = system
    N, ref_idx, radius, k = s["N"], 600, 0.6, 3000.0
    ctx = make_context(s)
    ctx.setup_local_md(TEMPERATURE, False)
    xs, _ = ctx.multiple_steps_local(1, np.array([ref_idx], dtype=np.int32), radius=radius, k=k, seed=21)
    free = ctx.local_md_free_idxs()
    frozen = np.setdiff1d(np.arange(N), free)
    assert ref_idx in free and len(frozen) > 100
    np.testing.assert_array_equal(xs[0][frozen], s["x"][frozen])
    fixed = np.zeros((N, 3), dtype=np.int64)
    for bp in make_context(s).get_potentials():
        fixed += np.rint(bp.execute(s["x"], s["box"])[0] * 2.0**36).astype(np.int64)
    others = free[free != ref_idx]
    fb = P.FlatBottomBond(np.stack([np.full(len(others), ref_idx), others], 1).astype(np.int32)).to_gpu(np.float32).unbound_impl
    fixed += np.rint(fb.execute(s["x"], np.tile([k, 0.0, radius], (len(others), 1)), s["box"])[0] * 2.0**36).astype(np.int64)
    beta = 1.0 / (0.008314462618 * TEMPERATURE)
    lfb = P.LogFlatBottomBond(np.stack([np.full(len(frozen), ref_idx), frozen], 1).astype(np.int32), beta).to_gpu(np.float32).unbound_impl
    fixed += np.rint(lfb.execute(s["x"], np.tile([k, 0.0, radius], (len(frozen), 1)), s["box"])[0] * 2.0**36).astype(np.int64)
    x1, v1 = O.baoab_step_mixed(s["x"], s["v"], fixed.view(np.uint64), s["masses"], TEMPERATURE, DT, 0.0, np.zeros((N, 3), np.float32))
    np.testing.assert_array_equal(xs[0][free], x1[free])

    ref = load_reference_ops()
    if ref is not None:
        ours, theirs = make_context(s), make_context(s, module=ref)
        ours.setup_local_md(TEMPERATURE, False)
        theirs.setup_local_md(TEMPERATURE, False)
        local_idxs = np.array([ref_idx, 77], dtype=np.int32)
        a, _ = ours.multiple_steps_local(20, local_idxs, radius=radius, k=k, seed=8)
        b, _ = theirs.multiple_steps_local(20, local_idxs, 0, radius, k, 8)
        np.testing.assert_array_equal(np.any(a[-1] != s["x"], axis=1), np.any(b[-1] != s["x"], axis=1))
        np.testing.assert_array_equal(a, b)  # including the log flat-bottom restraint on the frozen shell


def test_log_flat_bottom_bond_potential():
    """LogFlatBottomBond_{f32,f64} against the oracle (pinned to the reference's Python potential) and, where built, the
    compiled reference; argument checks with the reference's messages."""
    from pathlib import Path

    ops, lib, P = mods()
    g = np.load(Path(__file__).parent / "golden" / "log_flat_bottom_bond.npz")
    x, box, idxs, params, beta = g["x"], g["box"], g["idxs"], g["params"], float(g["beta"])
    ref_u, ref_dx, ref_dp = O.log_flat_bottom_bond(x, params, box, idxs, beta)
    np.testing.assert_allclose(ref_u, g["u"], rtol=1e-12)
    for precision, rtol in ((np.float64, 1e-9), (np.float32, 2e-4)):
        impl = P.LogFlatBottomBond(idxs, beta).to_gpu(precision).unbound_impl
        xx, pp = (x, params) if precision == np.float64 else (round_to_f32(x), round_to_f32(params))
        ou, odx, odp = O.log_flat_bottom_bond(xx, pp, box, idxs, beta)
        dx, dp, u = impl.execute(xx, pp, box)
        np.testing.assert_allclose(u, ou, rtol=rtol)
        np.testing.assert_allclose(dx, odx, rtol=rtol, atol=rtol * np.abs(odx).max())
        np.testing.assert_allclose(dp, odp, rtol=rtol, atol=rtol * np.abs(odp).max())
        dx2, dp2, u2 = impl.execute(xx, pp, box)
        np.testing.assert_array_equal(dx, dx2)
        assert u == u2
        assert impl.execute(xx, pp, box, True, False, False)[2] is None
    ref = load_reference_ops()
    if ref is not None:
        rimpl = ref.LogFlatBottomBond_f32(idxs, beta)
        rdx, rdp, ru = rimpl.execute(round_to_f32(x), round_to_f32(params), box, True, True, True)
        dx, dp, u = P.LogFlatBottomBond(idxs, beta).to_gpu(np.float32).unbound_impl.execute(round_to_f32(x), round_to_f32(params), box)
        # the reference's operation sequence, including its f64 evaluation of the two transcendental expressions: bitwise
        assert u == ru
        np.testing.assert_array_equal(dx, rdx)
        np.testing.assert_array_equal(dp, rdp)
    with pytest.raises(RuntimeError, match="beta must be positive"):
        ops.LogFlatBottomBond_f32(idxs, 0.0)
    with pytest.raises(RuntimeError, match=r"bond_idxs.size\(\) must be exactly 2\*k!"):
        ops.LogFlatBottomBond_f32(np.array([0, 1, 2], dtype=np.int32), 1.0)
    with pytest.raises(RuntimeError, match="src == dst"):
        ops.LogFlatBottomBond_f32(np.array([[1, 1]], dtype=np.int32), 1.0)
    with pytest.raises(RuntimeError, match=r"LogFlatBottomBond::execute_device\(\): expected P == 90, got P=3"):
        ops.LogFlatBottomBond_f32(idxs, 1.0).execute(x, params[:1], box)
