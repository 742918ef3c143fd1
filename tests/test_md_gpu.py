"""GPU: LangevinIntegrator + Context (reference tests/test_md.py): deterministic trajectory vs a NumPy BAOAB loop on
oracle forces (friction = 0, as the reference does because the noise streams cannot match), bit-exact integrator
arithmetic with injected noise, frame-store semantics, CUDA-graph replay == eager stepping, box validation errors,
noise statistics of the in-kernel Philox generator."""

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import round_to_f32, water_box

pytestmark = pytest.mark.gpu
BETA, CUTOFF = 2.0, 1.2


def ops():
    from timemachine_b200 import custom_ops

    return custom_ops


def pots():
    from timemachine_b200 import potentials

    return potentials


def small_bonded_system(rng, n=8):
    x = round_to_f32(rng.normal(0, 0.15, (n, 3)) + np.arange(n)[:, None] * 0.12)
    bond_idxs = np.array([(i, i + 1) for i in range(n - 1)], dtype=np.int32)
    bond_params = round_to_f32(np.stack([rng.uniform(2000, 5000, n - 1), rng.uniform(0.1, 0.15, n - 1)], 1))
    angle_idxs = np.array([(i, i + 1, i + 2) for i in range(n - 2)], dtype=np.int32)
    angle_params = round_to_f32(np.stack([rng.uniform(100, 300, n - 2), rng.uniform(1.5, 2.2, n - 2), np.zeros(n - 2)], 1))
    return x, bond_idxs, bond_params, angle_idxs, angle_params


@pytest.mark.parametrize("use_graphs", [False, True])
def test_fwd_mode_matches_numpy_baoab(rng, use_graphs):
    """tests/test_md.py:142-247: friction = 0 so cc = 0; Context trajectory vs a NumPy loop with oracle forces."""
    n = 8
    x0, bond_idxs, bond_params, angle_idxs, angle_params = small_bonded_system(rng, n)
    v0 = rng.normal(0, 0.5, (n, 3))
    masses = rng.uniform(1.0, 12.0, n)
    box = np.eye(3) * 100.0
    dt, temperature, friction = 1.5e-3, 300.0, 0.0
    o = ops()
    bp1 = o.BoundPotential(o.HarmonicBond_f64(bond_idxs), bond_params)
    bp2 = o.BoundPotential(o.HarmonicAngle_f64(angle_idxs), angle_params)
    intg = o.LangevinIntegrator(masses, temperature, dt, friction, 2022)
    ctx = o.Context(x0, v0, box, intg, [bp1, bp2])
    ctx.set_use_graphs(use_graphs)
    n_steps = 45
    xs, boxes = ctx.multiple_steps(n_steps, 5)
    assert xs.shape == (9, n, 3) and boxes.shape == (9, 3, 3)

    ca, cb, cc = O.langevin_coefficients(temperature, dt, friction, masses)
    assert np.all(cc == 0)
    x, v = x0.copy(), v0.copy()
    ref_frames = []
    for step in range(1, n_steps + 1):
        f = -(O.harmonic_bond(x, bond_params, bond_idxs)[1] + O.harmonic_angle(x, angle_params, angle_idxs)[1])
        x, v = O.baoab_step(x, v, f, ca, cb, cc, dt, np.zeros_like(x))
        if step % 5 == 0:
            ref_frames.append(x.copy())
    # f32 integrator arithmetic on f64 state: agreement at the 1e-6 level over 45 steps
    np.testing.assert_allclose(xs, np.array(ref_frames), rtol=0, atol=2e-6)
    np.testing.assert_allclose(ctx.get_x_t(), x, rtol=0, atol=2e-6)
    np.testing.assert_allclose(ctx.get_v_t(), v, rtol=0, atol=2e-4)
    np.testing.assert_array_equal(ctx.get_box(), box)


def test_integrator_arithmetic_is_bit_exact_with_injected_noise(rng):
    """One step with known fixed-point forces and known noise reproduces k_integrator.cuh:32-46 bit for bit
    (oracle.baoab_step_mixed is the NumPy statement of that casting sequence)."""
    n = 64
    o = ops()
    x0 = rng.normal(0, 1.0, (n, 3))
    v0 = rng.normal(0, 0.7, (n, 3))
    masses = rng.uniform(1.0, 16.0, n)
    box = np.eye(3) * 50.0
    temperature, dt, friction = 300.0, 2.5e-3, 1.0
    # a potential whose fixed-point du_dx we can read back exactly
    bond_idxs = np.array([(i, i + 1) for i in range(n - 1)], dtype=np.int32)
    bond_params = np.stack([rng.uniform(100, 1000, n - 1), rng.uniform(0.5, 1.5, n - 1)], 1)
    pot = o.HarmonicBond_f32(bond_idxs)
    du_dx = pot.execute(x0, bond_params, box, True, False, False)[0]
    fixed = np.rint(du_dx * 2.0**36).astype(np.int64).view(np.uint64)  # exact: du_dx is fixed/2^36
    noise = rng.normal(0, 1, (n, 3)).astype(np.float32)
    intg = o.LangevinIntegrator(masses, temperature, dt, friction, 7)
    intg.set_noise(noise)
    ctx = o.Context(x0, v0, box, intg, [o.BoundPotential(pot, bond_params)])
    ctx.step()
    ref_x, ref_v = O.baoab_step_mixed(x0, v0, fixed, masses, temperature, dt, friction, noise)
    np.testing.assert_array_equal(ctx.get_v_t(), ref_v)
    np.testing.assert_array_equal(ctx.get_x_t(), ref_x)


def water_context(n_waters, seed, precision=np.float32, friction=1.0, dt=1.0e-3, rng_seed=11, padding=0.1):
    s = water_box(n_waters, seed=seed)
    P = pots()
    nb = P.Nonbonded(s["N"], s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF, nblist_padding=padding)
    pot = P.SummedPotential(
        [P.HarmonicBond(s["bond_idxs"]), P.HarmonicAngle(s["angle_idxs"]), nb], [s["bond_params"], s["angle_params"], s["params"]]
    )
    flat = np.concatenate([s["bond_params"].reshape(-1), s["angle_params"].reshape(-1), s["params"].reshape(-1)])
    impl = pot.to_gpu(precision).unbound_impl
    o = ops()
    bp = o.BoundPotential(impl, flat)
    intg = o.LangevinIntegrator(s["masses"], 300.0, dt, friction, rng_seed)
    x0 = round_to_f32(s["x"])
    v0 = np.zeros_like(x0)
    return o.Context(x0, v0, s["box"], intg, [bp]), s, impl, flat


def test_graph_replay_equals_eager_stepping():
    """CUDA-graph replay, including device-side neighbour-list rebuilds and the Hilbert re-sort cadence, must give
    the same trajectory bit for bit as launching every kernel eagerly (same seed => same Philox noise)."""
    a, s, _, _ = water_context(700, seed=2)
    b, _, _, _ = water_context(700, seed=2)
    a.set_use_graphs(True)
    b.set_use_graphs(False)
    xa, _ = a.multiple_steps(237, 79)
    xb, _ = b.multiple_steps(237, 79)
    np.testing.assert_array_equal(xa, xb)
    np.testing.assert_array_equal(a.get_v_t(), b.get_v_t())
    # and one long call equals several short ones
    c, _, _, _ = water_context(700, seed=2)
    for k in (7, 100, 1, 50, 79):
        c.multiple_steps(k, k + 1)
    np.testing.assert_array_equal(c.get_x_t(), a.get_x_t())


def test_md_is_stable_and_thermalises():
    ctx, s, _, _ = water_context(700, seed=5, dt=1.5e-3)
    ctx.multiple_steps(1500, 2000)  # no frames
    v = ctx.get_v_t()
    x = ctx.get_x_t()
    assert np.isfinite(x).all() and np.isfinite(v).all()
    ke = 0.5 * np.sum(s["masses"][:, None] * v * v)
    T = 2 * ke / (3 * s["N"] * O.BOLTZ)
    assert 200.0 < T < 400.0, T
    # bond lengths stay physical
    d = np.linalg.norm(x[s["bond_idxs"][:, 0]] - x[s["bond_idxs"][:, 1]], axis=1)
    assert d.max() < 0.13 and d.min() > 0.07


def test_store_x_interval_semantics():
    """tests/test_md.py:24-76"""
    ctx, s, _, _ = water_context(700, seed=6)
    N = s["N"]
    xs, boxes = ctx.multiple_steps(10)
    assert xs.shape == (1, N, 3) and boxes.shape == (1, 3, 3)
    np.testing.assert_array_equal(xs[0], ctx.get_x_t())
    xs, boxes = ctx.multiple_steps(10, 10)
    assert xs.shape == (1, N, 3)
    xs, boxes = ctx.multiple_steps(10, 11)
    assert xs.shape == (0, N, 3) and boxes.shape == (0, 3, 3)
    xs, boxes = ctx.multiple_steps(10, 3)
    assert xs.shape == (3, N, 3)
    xs, boxes = ctx.multiple_steps(10, 1)
    assert xs.shape == (10, N, 3)
    np.testing.assert_array_equal(xs[-1], ctx.get_x_t())
    with pytest.raises(RuntimeError, match="store_x_interval must be greater than or equal to zero"):
        ctx.multiple_steps(10, -1)


def test_set_get_state():
    ctx, s, _, _ = water_context(700, seed=6)
    x = ctx.get_x_t() + 0.001
    v = np.full_like(x, 0.25)
    box = s["box"] * 1.01
    ctx.set_x_t(x)
    ctx.set_v_t(v)
    ctx.set_box(box)
    np.testing.assert_array_equal(ctx.get_x_t(), x)
    np.testing.assert_array_equal(ctx.get_v_t(), v)
    np.testing.assert_array_equal(ctx.get_box(), box)
    with pytest.raises(RuntimeError):
        ctx.set_x_t(x[:-1])
    assert ctx.get_integrator() is not None and len(ctx.get_potentials()) == 1 and ctx.get_movers() == []


def test_box_validation_errors():
    """tests/test_md.py:895-976"""
    o = ops()
    s = water_box(150, seed=6)
    intg = o.LangevinIntegrator(s["masses"], 300.0, 1e-3, 1.0, 1)
    x = s["x"]
    with pytest.raises(RuntimeError, match="box must be 3x3"):
        o.Context(x, np.zeros_like(x), np.eye(2), intg, [])
    with pytest.raises(RuntimeError, match="box must have positive values along diagonal"):
        o.Context(x, np.zeros_like(x), np.zeros((3, 3)), intg, [])
    bad = s["box"].copy()
    bad[0, 1] = 0.1
    with pytest.raises(RuntimeError, match="box must be ortholinear"):
        o.Context(x, np.zeros_like(x), bad, intg, [])
    with pytest.raises(RuntimeError, match="v0 N != x0 N"):
        o.Context(x, np.zeros((3, 3)), s["box"], intg, [])
    # a box smaller than 2 (cutoff + padding) is rejected when a frame is collected
    ctx, s2, _, _ = water_context(150, seed=6)
    ctx.set_box(np.eye(3) * 2.0)
    with pytest.raises(RuntimeError, match="cutoff with padding is more than half of the box width"):
        ctx.multiple_steps(2)
    # ... but not when no frame is collected (store_x_interval > n_steps skips validation)
    ctx2, _, _, _ = water_context(150, seed=6)
    ctx2.set_box(np.eye(3) * 2.0)
    ctx2.multiple_steps(2, 3)


def test_philox_noise_statistics():
    o = ops()
    n = 200_000
    a = o.fill_normal(n, seed=123, step=0).astype(np.float64)
    assert abs(a.mean()) < 5e-3 and abs(a.std() - 1.0) < 5e-3
    assert abs(np.mean(a**3)) < 2e-2 and abs(np.mean(a**4) - 3.0) < 5e-2
    # components and atoms are uncorrelated; different steps / seeds give different, reproducible streams
    assert abs(np.corrcoef(a[:, 0], a[:, 1])[0, 1]) < 1e-2 and abs(np.corrcoef(a[:-1, 0], a[1:, 0])[0, 1]) < 1e-2
    b = o.fill_normal(n, seed=123, step=1)
    c = o.fill_normal(n, seed=124, step=0)
    assert abs(np.corrcoef(a[:, 0], b[:, 0])[0, 1]) < 1e-2 and abs(np.corrcoef(a[:, 0], c[:, 0])[0, 1]) < 1e-2
    np.testing.assert_array_equal(o.fill_normal(1000, 123, 0), a[:1000].astype(np.float32))
    assert np.abs(a).max() < 7.0


def test_captured_graph_is_dropped_when_launch_state_changes():
    """Round-1 advisor finding: the step graph bakes in kernel arguments (atom counts and grid sizes of the all-pairs
    potential, the integrator's noise pointer).  Mutating them between two multiple_steps calls must not replay the stale
    graph: after set_atom_idxs / set_noise the graph context equals, bit for bit, an eager context that went through the
    same calls."""
    a, s, impl_a, _ = water_context(700, seed=4)
    b, _, impl_b, _ = water_context(700, seed=4)
    a.set_use_graphs(True)
    b.set_use_graphs(False)
    N = s["N"]
    for ctx in (a, b):
        ctx.multiple_steps(60, 61)  # captures (a) the 10-step graph
    np.testing.assert_array_equal(a.get_x_t(), b.get_x_t())
    # restrict the all-pairs term to two thirds of the atoms: K, NR and every grid derived from them change
    subset = np.arange(0, 2 * N // 3, dtype=np.int32)
    for impl in (impl_a, impl_b):
        impl.get_potentials()[2].get_potentials()[0].set_atom_idxs(subset)
    for ctx in (a, b):
        ctx.multiple_steps(40, 41)
    np.testing.assert_array_equal(a.get_x_t(), b.get_x_t())
    np.testing.assert_array_equal(a.get_v_t(), b.get_v_t())
    # the forces really changed (the comparison above is not vacuous): a third context that keeps all atoms differs
    c, _, _, _ = water_context(700, seed=4)
    c.multiple_steps(100, 101)
    assert np.abs(c.get_x_t() - a.get_x_t()).max() > 1e-6
    # external noise switched on after a capture
    noise = np.random.default_rng(1).normal(size=(N, 3)).astype(np.float32)
    for ctx in (a, b):
        ctx.get_integrator().set_noise(noise)
        ctx.multiple_steps(30, 31)
    np.testing.assert_array_equal(a.get_x_t(), b.get_x_t())
    for ctx in (a, b):
        ctx.get_integrator().set_noise(None)
        ctx.multiple_steps(30, 31)
    np.testing.assert_array_equal(a.get_x_t(), b.get_x_t())
    np.testing.assert_array_equal(a.get_v_t(), b.get_v_t())
