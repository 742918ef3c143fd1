"""GPU: MonteCarloBarostat, modelled on the reference's tests/test_barostat.py:
argument validation with the reference's messages, the move against the CPU oracle given the same uniforms, bit-for-bit
agreement with the compiled reference over a sequence of moves with the same seed (same cuRAND stream), determinism,
partial group lists, response to pressure, and the Context integration (NPT run, CUDA-graph replay == eager)."""

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import load_reference_ops, round_to_f32, water_box

pytestmark = pytest.mark.gpu
BETA, CUTOFF = 2.0, 1.2
TEMPERATURE, PRESSURE = 300.0, 1.013


def ops():
    from timemachine_b200 import custom_ops

    return custom_ops


def make_bps(o, w, precision="f32"):
    """The water-box potentials as separate bound potentials, reference-style (bond, angle, all-pairs, exclusions)."""
    N = w["N"]
    return [
        o.BoundPotential(getattr(o, f"HarmonicBond_{precision}")(w["bond_idxs"]), w["bond_params"]),
        o.BoundPotential(getattr(o, f"HarmonicAngle_{precision}")(w["angle_idxs"]), w["angle_params"]),
        o.BoundPotential(getattr(o, f"NonbondedAllPairs_{precision}")(N, BETA, CUTOFF), w["params"]),
        o.BoundPotential(
            getattr(o, f"NonbondedExclusions_{precision}")(w["exclusion_idxs"], w["scale_factors"], BETA, CUTOFF), w["params"]
        ),
    ]


def water_groups(w):
    return [np.arange(i, i + 3, dtype=np.int32) for i in range(0, w["N"], 3)]


@pytest.fixture(scope="module")
def water():
    w = water_box(700, seed=11)
    w["x"] = round_to_f32(w["x"])
    w["params"] = round_to_f32(w["params"])
    return w


def total_energy_fixed(bps, x, box):
    return sum(int(np.asarray(bp.execute_fixed(x, box)).astype(np.int64)[0]) for bp in bps)


# ---- validation (tests/test_barostat.py:21-70, 137-176) ----------------------------------------------------------------
def test_barostat_validation(water):
    o = ops()
    bps = make_bps(o, water)
    N, groups = water["N"], water_groups(water)
    with pytest.raises(RuntimeError, match="interval must be greater than 0"):
        o.MonteCarloBarostat(N, PRESSURE, TEMPERATURE, groups, 0, bps, 1, True, 0.0)
    with pytest.raises(RuntimeError, match="interval must be greater than 0"):
        o.MonteCarloBarostat(N, PRESSURE, TEMPERATURE, groups, -1, bps, 1, True, 0.0)
    with pytest.raises(RuntimeError, match="Grouped indices must be between 0 and N"):
        o.MonteCarloBarostat(N, PRESSURE, TEMPERATURE, [[0, 1, N]], 5, bps, 1, True, 0.0)
    with pytest.raises(RuntimeError, match="Grouped indices must be between 0 and N"):
        o.MonteCarloBarostat(N, PRESSURE, TEMPERATURE, [[-1, 1, 2]], 5, bps, 1, True, 0.0)
    with pytest.raises(RuntimeError, match="All grouped indices must be unique"):
        o.MonteCarloBarostat(N, PRESSURE, TEMPERATURE, [[0, 1, 2], [2, 3]], 5, bps, 1, True, 0.0)
    baro = o.MonteCarloBarostat(N, PRESSURE, TEMPERATURE, groups, 5, bps, 1, True, 0.0)
    assert baro.get_interval() == 5 and baro.get_adaptive_scaling() and baro.get_volume_scale_factor() == 0.0
    with pytest.raises(RuntimeError, match="interval must be greater than 0"):
        baro.set_interval(0)
    with pytest.raises(RuntimeError, match="step must be at least 0"):
        baro.set_step(-1)
    baro.set_interval(3)
    assert baro.get_interval() == 3
    baro.set_adaptive_scaling(False)
    assert not baro.get_adaptive_scaling()
    baro.set_volume_scale_factor(0.75)
    assert baro.get_volume_scale_factor() == 0.75
    with pytest.raises(RuntimeError, match="N != N_"):
        baro.move(water["x"][:-3], water["box"])


def test_barostat_acts_every_interval_calls(water):
    o = ops()
    bps = make_bps(o, water)
    baro = o.MonteCarloBarostat(water["N"], PRESSURE, TEMPERATURE, water_groups(water), 3, bps, 7, True, 0.0)
    x, box = water["x"], water["box"]
    for call in range(1, 10):
        x2, box2 = baro.move(x, box)
        attempted = sum(baro.counters()[:1])
        assert attempted == call // 3  # counters are only reset by the adaptive rule after 10 attempts
        if call % 3 != 0:
            np.testing.assert_array_equal(x2, x)
            np.testing.assert_array_equal(box2, box)
    baro.set_step(2)  # the next call is the 3rd
    baro.move(x, box)
    assert baro.counters()[0] == 4


# ---- the move against the oracle -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_move_matches_oracle(water, seed):
    o = ops()
    bps = make_bps(o, water)
    groups = water_groups(water)
    N, x, box = water["N"], water["x"], water["box"]
    baro = o.MonteCarloBarostat(N, PRESSURE, TEMPERATURE, groups, 1, bps, seed, True, 0.0)
    x_new, box_new = baro.move(x, box)
    rand0, rand1 = baro.last_uniforms()
    assert 0.0 < rand0 <= 1.0 and 0.0 < rand1 <= 1.0
    x_prop, box_prop, volume, delta, scale_used = O.barostat_propose(x, box, groups, 0.0, rand0, adaptive=True)
    np.testing.assert_allclose(baro.get_volume_scale_factor(), scale_used, rtol=1e-12)
    u0 = total_energy_fixed(bps, x, box)
    u1 = total_energy_fixed(bps, x_prop, box_prop)
    accepted, w = O.barostat_accepts(u0, u1, volume, delta, len(groups), TEMPERATURE, PRESSURE, rand1)
    assert baro.counters() == (1, 1 if accepted else 0), f"w={w} rand1={rand1}"
    if accepted:
        # cbrtf differs by an ulp between libm and libdevice: the scale, hence every coordinate, agrees to ~1e-6
        np.testing.assert_allclose(np.diag(box_new), np.diag(box_prop), rtol=3e-7)
        np.testing.assert_allclose(x_new, x_prop, rtol=0, atol=5e-6)
    else:
        np.testing.assert_array_equal(x_new, x)
        np.testing.assert_array_equal(box_new, box)


def test_accepted_move_is_rigid_and_in_the_home_box(water):
    o = ops()
    # no potentials: dU == 0; a compression at 1 bar changes w by ~ -N kT ln(V'/V) > 0 ... use many moves and keep the accepted
    groups = water_groups(water)
    N, x, box = water["N"], water["x"], water["box"]
    baro = o.MonteCarloBarostat(N, PRESSURE, TEMPERATURE, groups, 1, [], 5, False, 0.5)
    seen = 0
    for _ in range(20):
        x2, box2 = baro.move(x, box)
        if np.array_equal(box2, box):
            continue
        seen += 1
        scale = box2[0, 0] / box[0, 0]
        np.testing.assert_allclose(np.diag(box2) / np.diag(box), scale, rtol=1e-12)
        xm, xm2 = x.reshape(-1, 3, 3), x2.reshape(-1, 3, 3)
        np.testing.assert_allclose(xm2 - xm2[:, :1], xm - xm[:, :1], rtol=0, atol=1e-12)  # rigid molecules
        cent = xm2.mean(axis=1)
        assert np.all(cent > -1e-5) and np.all(cent < np.diag(box2) + 1e-5)
    assert seen >= 5


def test_partial_group_idxs_leave_other_atoms_alone(water):
    """tests/test_barostat.py:179-240: atoms outside every group are not scaled."""
    o = ops()
    groups = water_groups(water)[:100]
    N, x, box = water["N"], water["x"], water["box"]
    baro = o.MonteCarloBarostat(N, PRESSURE, TEMPERATURE, groups, 1, [], 3, False, 0.5)
    moved = False
    for _ in range(20):
        x2, box2 = baro.move(x, box)
        np.testing.assert_array_equal(x2[300:], x[300:])
        if not np.array_equal(box2, box):
            moved = True
            assert not np.array_equal(x2[:300], x[:300])
    assert moved


# ---- bit-for-bit against the compiled reference -------------------------------------------------------------------------
@pytest.mark.parametrize("seed,adaptive,initial_scale", [(2022, True, 0.0), (5, False, 0.4)])
def test_sequence_of_moves_equals_reference_bitwise(water, seed, adaptive, initial_scale):
    ref = load_reference_ops()
    if ref is None:
        pytest.skip("oracle/_ref/custom_ops*.so not built")
    o = ops()
    groups = water_groups(water)
    N = water["N"]
    ours = o.MonteCarloBarostat(N, PRESSURE, TEMPERATURE, groups, 1, make_bps(o, water), seed, adaptive, initial_scale)
    theirs = ref.MonteCarloBarostat(
        N, PRESSURE, TEMPERATURE, [g.tolist() for g in groups], 1, make_bps(ref, water), seed, adaptive, initial_scale
    )
    x, box = water["x"].copy(), water["box"].copy()
    xr, boxr = x.copy(), box.copy()
    n_changed = 0
    for it in range(40):
        x_new, box_new = ours.move(x, box)
        xr_new, boxr_new = theirs.move(xr, boxr)
        assert np.array_equal(box_new, boxr_new), f"move {it}: box differs {np.diag(box_new)} vs {np.diag(boxr_new)}"
        bad = np.argwhere(x_new != xr_new)
        assert len(bad) == 0, f"move {it}: {len(bad)} coordinates differ, first {bad[:3].tolist()}"
        assert ours.get_volume_scale_factor() == theirs.get_volume_scale_factor()
        n_changed += not np.array_equal(box_new, box)
        x, box, xr, boxr = x_new, box_new, xr_new, boxr_new
    assert 5 <= n_changed <= 40


# ---- determinism, pressure response ---------------------------------------------------------------------------------------
def test_barostat_is_deterministic(water):
    o = ops()
    groups = water_groups(water)

    def run(seed):
        baro = o.MonteCarloBarostat(water["N"], PRESSURE, TEMPERATURE, groups, 1, make_bps(o, water), seed, True, 0.0)
        x, box = water["x"], water["box"]
        for _ in range(15):
            x, box = baro.move(x, box)
        return x, box, baro.get_volume_scale_factor()

    a, b, c = run(42), run(42), run(43)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    assert a[2] == b[2]
    assert not np.array_equal(a[1], c[1])


def test_npt_run_and_pressure_response(water):
    """tests/test_barostat.py:305-357: a much higher pressure gives a smaller box; the run is reproducible, and CUDA-graph
    replay of the plain steps between barostat moves gives the eager trajectory bit for bit."""
    o = ops()
    groups = water_groups(water)
    N = water["N"]
    v0 = np.zeros((N, 3))

    def run(pressure, use_graphs, n_steps=300):
        bps = make_bps(o, water)
        intg = o.LangevinIntegrator(water["masses"], TEMPERATURE, 1.5e-3, 1.0, 2024)
        baro = o.MonteCarloBarostat(N, pressure, TEMPERATURE, groups, 15, bps, 2025, True, 0.0)
        ctx = o.Context(water["x"], v0, water["box"], intg, bps, movers=[baro])
        assert ctx.get_barostat() is baro and ctx.get_movers() == [baro]
        ctx.set_use_graphs(use_graphs)
        xs, boxes = ctx.multiple_steps(n_steps, 10)
        attempted = baro.counters()
        return xs, boxes, ctx.get_x_t(), attempted

    xs_g, boxes_g, x_g, _ = run(PRESSURE, True)
    xs_e, boxes_e, x_e, _ = run(PRESSURE, False)
    np.testing.assert_array_equal(boxes_g, boxes_e)
    np.testing.assert_array_equal(x_g, x_e)
    np.testing.assert_array_equal(xs_g, xs_e)
    vol = np.prod(np.diagonal(boxes_g, axis1=1, axis2=2), axis=1)
    assert np.all(np.isfinite(x_g)) and len(np.unique(vol)) > 1  # the box did change
    # the jittered-lattice start is far from equilibrium (dU dominates at ordinary pressures), so use a P dV term that cannot be missed
    _, boxes_hi, _, _ = run(1.0e5, True)
    vol_hi = np.prod(np.diagonal(boxes_hi, axis1=1, axis2=2), axis=1)
    assert vol_hi[-1] < vol[-1]
