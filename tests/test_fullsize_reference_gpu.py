"""GPU: BASELINE.json's configurations at FULL size against the unmodified reference custom_ops compiled for sm_100a
(oracle/_ref; mandatory on a GPU box, tests/conftest.py sets TMB_REQUIRE_REF=1).

  B  DHFR: the reference's own benchmark geometry (23,558 atoms, 6.223 nm; tests/golden/dhfr_5dfr.npz, topology derived
     by tests/dhfr_system.py), Summed[bond, angle, proper, improper, Nonbonded]: forces of every term and of the sum,
     du/dp, energy; a 100-step friction-0 trajectory of both implementations
  C  the bench's own leg (bench.build_system: 29,862 atoms, six potentials, lambda = 0.5) with du/dp
  E  90k-atom water box: nonbonded forces bitwise

Stated bar (north_star): forces within 1e-5 relative of the reference custom_ops.  What holds for every f32 kernel on the
path - the nonbonded tile and pair-list kernels, bonds, angles, torsions - is stronger: they perform the same sequence of
rounded operations per term as the reference's (the FMA contraction nvcc applies to the reference's source is written out
explicitly here, read off the SASS of oracle/_ref) and all sums are integers, so forces, du/dp and energies are BIT FOR BIT
equal (wherever no single pair term exceeds the int64 fixed-point range)."""

import numpy as np
import pytest

from tests.common import assert_forces_close, load_reference_ops, round_to_f32, water_box

pytestmark = pytest.mark.gpu
BETA, CUTOFF = 2.0, 1.2


def mods():
    from timemachine_b200 import custom_ops, potentials

    return custom_ops, potentials


def require_ref():
    ref = load_reference_ops()
    if ref is None:
        pytest.skip("oracle/_ref/custom_ops*.so not built")
    return ref


@pytest.fixture(scope="module")
def dhfr():
    from tests.dhfr_system import load_dhfr

    s = load_dhfr()
    for k in ("bond_params", "angle_params", "proper_params", "improper_params", "params", "scale_factors"):
        s[k] = round_to_f32(s[k])
    return s


def dhfr_terms(s, module, suffix):
    """[(name, potential, params)] on `module` (this repo's custom_ops or the compiled reference), same constructors."""
    N = s["N"]
    g = lambda name: getattr(module, f"{name}_{suffix}")  # noqa: E731
    nb = module.FanoutSummedPotential(
        [g("NonbondedAllPairs")(N, BETA, CUTOFF, None, False, 0.1), g("NonbondedExclusions")(s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF)], True
    )
    return [
        ("HarmonicBond", g("HarmonicBond")(s["bond_idxs"]), s["bond_params"]),
        ("HarmonicAngle", g("HarmonicAngle")(s["angle_idxs"]), s["angle_params"]),
        ("PeriodicTorsion(proper)", g("PeriodicTorsion")(s["proper_idxs"]), s["proper_params"]),
        ("PeriodicTorsion(improper)", g("PeriodicTorsion")(s["improper_idxs"]), s["improper_params"]),
        ("Nonbonded", nb, s["params"]),
    ]


def test_config_b_dhfr_every_term_and_the_sum_against_the_reference(dhfr):
    ref = require_ref()
    ops, _ = mods()
    s = dhfr
    x, box = s["x"], s["box"]
    mine, theirs, exact = dhfr_terms(s, ops, "f32"), dhfr_terms(s, ref, "f32"), dhfr_terms(s, ref, "f64")
    for (name, a, p), (_, b, _), (_, b64, _) in zip(mine, theirs, exact):
        dx, dp, u = a.execute(x, p, box)
        rdx, rdp, ru = b.execute(x, p, box)
        np.testing.assert_allclose(u, ru, rtol=1e-5, err_msg=name)
        np.testing.assert_allclose(dp, rdp, rtol=1e-4, atol=1e-4 * max(1.0, np.abs(rdp).max()), err_msg=name)
        # same sequence of rounded operations per term, integer sums: nothing may differ
        np.testing.assert_array_equal(dx, rdx, err_msg=name)
        np.testing.assert_array_equal(dp, rdp, err_msg=name)
        assert u == ru, name
        # and the f32 kernels sit where f32 round-off puts them relative to the reference's f64 kernels
        xdx, _, _ = b64.execute(x, p, box)
        assert_forces_close(xdx, dx, 5e-3, what=name + " vs the reference's f64 kernel")
    sizes = [p.size for _, _, p in mine]
    flat = np.concatenate([p.reshape(-1) for _, _, p in mine])
    summed = ops.SummedPotential([a for _, a, _ in mine], sizes, True)
    rsummed = ref.SummedPotential([b for _, b, _ in theirs], sizes, True)
    dx, dp, u = summed.execute(x, flat, box)
    rdx, rdp, ru = rsummed.execute(x, flat, box)
    np.testing.assert_array_equal(dx, rdx, err_msg="SummedPotential du/dx")
    np.testing.assert_array_equal(dp, rdp, err_msg="SummedPotential du/dp")
    assert u == ru
    # Newton's third law, exactly, in fixed point: holds for the pair terms (a column atom receives the negated integer of
    # its row atom); angle and torsion terms round the force on each of their atoms separately, like the reference
    nb_dx = mine[-1][1].execute(x, mine[-1][2], box)[0]
    assert not np.rint(nb_dx * 2.0**36).astype(np.int64).sum(axis=0).any()


def test_config_b_dhfr_trajectory_against_the_reference(dhfr):
    """100 steps at friction 0 (the noise coefficient vanishes: tests/test_md.py:174-178) from the same x0, v0: frames every 20
    steps and the final velocities of the two Contexts are bitwise equal."""
    ref = require_ref()
    ops, _ = mods()
    s = dhfr
    N = s["N"]
    rng = np.random.default_rng(7)
    v0 = rng.normal(0, 1.0, (N, 3)) * np.sqrt(0.008314462618 * 300.0 / s["masses"])[:, None]
    outs = []
    for module in (ops, ref):
        terms = dhfr_terms(s, module, "f32")
        sizes = [p.size for _, _, p in terms]
        flat = np.concatenate([p.reshape(-1) for _, _, p in terms])
        bp = module.BoundPotential(module.SummedPotential([a for _, a, _ in terms], sizes, True), flat)
        intg = module.LangevinIntegrator(s["masses"], 300.0, 1.5e-3, 0.0, 2024)
        ctx = module.Context(s["x"], v0, s["box"], intg, [bp])
        xs, boxes = ctx.multiple_steps(100, 20)
        outs.append((np.asarray(xs), np.asarray(ctx.get_v_t())))
    (xa, va), (xb, vb) = outs
    assert np.isfinite(xa).all()
    moved = np.abs(xa[-1] - s["x"]).max()
    assert moved > 0.02, "the atoms did not move: the comparison would prove nothing"
    # identical forces (bit for bit, see above) through an identical integrator: the two trajectories do not separate at all
    np.testing.assert_array_equal(xa, xb)
    np.testing.assert_array_equal(va, vb)


def test_config_c_bench_leg_against_the_reference():
    """The exact system bench.py times (29,862 atoms, lambda = 0.5): du/dx, du/dp (du/dlambda enters through w and q of the
    dummy atoms) and u of the six-potential sum, and of the two tile-kernel potentials alone, against the reference."""
    import bench as B

    ref = require_ref()
    ops, P = mods()
    s = B.build_system(10000, 60, seed=2022)
    N = s["N"]
    flat = round_to_f32(B.flat_params(s, 0.5))
    x = round_to_f32(s["x"])
    impl = B.make_potential(P, s).to_gpu(np.float32).unbound_impl
    rimpl = B.make_reference_potential(ref, s)
    dx, dp, u = impl.execute(x, flat, s["box"])
    rdx, rdp, ru = rimpl.execute(x, flat, s["box"])
    np.testing.assert_array_equal(dx, rdx, err_msg="bench leg du/dx")
    np.testing.assert_array_equal(dp, rdp, err_msg="bench leg du/dp")
    assert u == ru
    # du/dp of the interaction group alone (where lambda enters): bitwise, all four columns
    p = round_to_f32(B.params_at_lambda(s, 0.5))
    ixn = ops.NonbondedInteractionGroup_f32(N, s["lig_idx"], BETA, CUTOFF, s["env_idx"], False, 0.1)
    rixn = ref.NonbondedInteractionGroup_f32(N, s["lig_idx"], BETA, CUTOFF, s["env_idx"], False, 0.1)
    a, b = ixn.execute(x, p, s["box"]), rixn.execute(x, p, s["box"])
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    assert a[2] == b[2]
    assert np.abs(a[1][s["dummy"], 3]).max() > 0  # du/dw of the decoupled atoms is non-trivial
    ap = ops.NonbondedAllPairs_f32(N, BETA, CUTOFF, s["env_idx"], False, 0.1)
    rap = ref.NonbondedAllPairs_f32(N, BETA, CUTOFF, s["env_idx"], False, 0.1)
    a, b = ap.execute(x, p, s["box"]), rap.execute(x, p, s["box"])
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    assert a[2] == b[2]


def test_config_e_90k_forces_bitwise_against_the_reference():
    ref = require_ref()
    ops, _ = mods()
    s = water_box(30000, seed=30000)
    N = s["N"]
    x, params, box = round_to_f32(s["x"]), round_to_f32(s["params"]), s["box"]
    mine = ops.FanoutSummedPotential(
        [ops.NonbondedAllPairs_f32(N, BETA, CUTOFF, None, False, 0.1), ops.NonbondedExclusions_f32(s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF)], True
    )
    theirs = ref.FanoutSummedPotential(
        [ref.NonbondedAllPairs_f32(N, BETA, CUTOFF, None, False, 0.1), ref.NonbondedExclusions_f32(s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF)], True
    )
    dx, dp, u = mine.execute(x, params, box)
    rdx, rdp, ru = theirs.execute(x, params, box)
    np.testing.assert_array_equal(dx, rdx)
    np.testing.assert_array_equal(dp, rdp)
    assert u == ru
