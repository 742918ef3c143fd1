"""GPU: the drop-in claim, second half.  The constructor calls the reference's UNMODIFIED wrappers make
(tests/golden/dropin_calls.{json,npz}, recorded from /root/reference by tests/golden/make_golden_dropin.py; the CPU half
in tests/test_dropin_cpu.py shows that this repo's mirror dataclasses make exactly the same calls) are replayed against the
real module: every object constructs, every potential evaluates, the f32 and f64 classes agree, the composite the
free-energy code builds matches the oracle, the integrator + barostat built from the recorded calls run in a Context.
Also here: the JVP rule of the reference's jax_interface (`timemachine_b200.jax_interface`) against finite differences."""

import json
from pathlib import Path

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import assert_forces_close, round_to_f32, water_box

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"

# parameters per term for the classes whose parameter array is per term; nonbonded classes take [N, 4]
PER_TERM = {"HarmonicBond": 2, "HarmonicAngle": 3, "PeriodicTorsion": 3, "FlatBottomBond": 3, "LogFlatBottomBond": 3,
            "ChiralAtomRestraint": 1, "ChiralBondRestraint": 1, "NonbondedPairListPrecomputed": 4}


def replay():
    from timemachine_b200 import custom_ops

    calls = json.loads((GOLDEN / "dropin_calls.json").read_text())["calls"]
    arrays = dict(np.load(GOLDEN / "dropin_calls.npz"))
    objs = {}

    def dec(v):
        if "obj" in v:
            return objs[v["obj"]]
        if "array" in v:
            return arrays[v["array"]]
        if "list" in v:
            return [dec(x) for x in v["list"]]
        if "none" in v:
            return None
        return next(iter(v.values()))

    for c in calls:
        objs[c["id"]] = getattr(custom_ops, c["cls"])(*[dec(a) for a in c["args"]], **{k: dec(v) for k, v in c["kwargs"].items()})
    return calls, arrays, objs


def test_replayed_reference_calls_construct_and_evaluate():
    calls, arrays, objs = replay()
    s = water_box(64, seed=5)
    x, box, N = round_to_f32(s["x"]), s["box"] * 1.0, s["N"]
    box = np.eye(3) * 3.0  # 2 (cutoff + padding) < 3 nm
    rng = np.random.default_rng(3)
    results = {}
    n_exec = 0
    for c in calls:
        base, _, suffix = c["cls"].rpartition("_")
        if suffix not in ("f32", "f64"):
            continue
        impl = objs[c["id"]]
        if base in PER_TERM:
            n_terms = len(arrays[c["args"][0]["array"]])
            width = PER_TERM[base]
            if base == "HarmonicBond":
                p = np.stack([np.full(n_terms, 1000.0), np.full(n_terms, 0.1)], 1)
            elif base == "HarmonicAngle":
                p = np.stack([np.full(n_terms, 100.0), np.full(n_terms, 1.8), np.zeros(n_terms)], 1)
            elif base == "PeriodicTorsion":
                p = np.stack([np.full(n_terms, 5.0), np.full(n_terms, 0.3), np.full(n_terms, 2.0)], 1)
            elif base in ("FlatBottomBond", "LogFlatBottomBond"):
                p = np.stack([np.full(n_terms, 500.0), np.full(n_terms, 0.3), np.full(n_terms, 0.2)], 1)
            elif base == "NonbondedPairListPrecomputed":
                p = np.stack([np.full(n_terms, 0.5), np.full(n_terms, 0.3), np.full(n_terms, 0.2), np.zeros(n_terms)], 1)
            else:
                p = np.full((n_terms, width), 50.0)
        elif base == "CentroidRestraint":
            p = np.zeros(0)  # takes no parameters
        else:
            p = s["params"]
        p = round_to_f32(p)
        du_dx, du_dp, u = impl.execute(x, p, box)
        assert np.isfinite(u) and np.isfinite(du_dx).all() and du_dp.shape == p.shape, c["cls"]
        again = impl.execute(x, p, box)
        np.testing.assert_array_equal(du_dx, again[0])
        # the wrappers construct the f32 and the f64 class of a potential one after the other: pair them by order
        slot = results.setdefault(base, {"f32": [], "f64": []})
        slot[suffix].append((du_dx, u))
        n_exec += 1
    assert n_exec >= 32
    # the two precisions of one recorded constructor call agree
    n_pairs = 0
    for base, v in results.items():
        for (dx32, u32), (dx64, u64) in zip(v["f32"], v["f64"]):
            assert_forces_close(dx64, dx32, 1e-3, what=base)
            np.testing.assert_allclose(u32, u64, rtol=2e-4, atol=1e-2, err_msg=base)
            n_pairs += 1
    assert n_pairs >= 15


def test_replayed_composite_matches_the_oracle_and_runs_md():
    from timemachine_b200 import custom_ops

    calls, arrays, objs = replay()
    s = water_box(64, seed=5)
    N = s["N"]
    box = np.eye(3) * 3.0
    x = round_to_f32(s["x"])
    summed = next(objs[c["id"]] for c in calls if c["cls"] == "SummedPotential")
    assert isinstance(summed, custom_ops.SummedPotential) and len(summed.get_potentials()) == 3
    flat = np.concatenate([round_to_f32(s["bond_params"]).reshape(-1), round_to_f32(s["angle_params"]).reshape(-1), round_to_f32(s["params"]).reshape(-1)])
    du_dx, du_dp, u = summed.execute(x, flat, box)
    ub, dxb, _ = O.harmonic_bond(x, round_to_f32(s["bond_params"]), s["bond_idxs"])
    ua, dxa, _ = O.harmonic_angle(x, round_to_f32(s["angle_params"]), s["angle_idxs"])
    un, dxn, _ = O.nonbonded(x, round_to_f32(s["params"]), box, s["exclusion_idxs"], s["scale_factors"], 2.0, 1.2)
    np.testing.assert_allclose(u, ub + ua + un, rtol=1e-4, atol=5e-3)
    assert_forces_close(dxb + dxa + dxn, du_dx, 1e-4)
    # BoundPotential / LangevinIntegrator / MonteCarloBarostat exactly as the reference's lib dataclasses built them
    bound = next(objs[c["id"]] for c in calls if c["cls"] == "BoundPotential" and "obj" in c["args"][0] and calls[c["args"][0]["obj"]]["cls"] == "SummedPotential")
    intg = next(objs[c["id"]] for c in calls if c["cls"] == "LangevinIntegrator")
    baro = next(objs[c["id"]] for c in calls if c["cls"] == "MonteCarloBarostat")
    np.testing.assert_array_equal(bound.execute(x, box)[0], du_dx)
    ctx = custom_ops.Context(x, np.zeros_like(x), box, intg, [bound], movers=[baro])
    xs, boxes = ctx.multiple_steps(60, 20)
    assert xs.shape == (3, N, 3) and np.isfinite(xs).all() and np.isfinite(boxes).all()
    assert baro.get_interval() == 15


@pytest.mark.parametrize("precision,tol", [(np.float64, 5e-3), (np.float32, 5e-3)])
def test_jvp_rule_against_finite_differences(precision, tol):
    """reference potentials/jax_interface.py:27-46 restated in NumPy (timemachine_b200/jax_interface.py): the directional
    derivative u'(0) of u(x + t dx, p + t dp) equals sum(du_dx dx) + sum(du_dp dp)."""
    from timemachine_b200 import jax_interface as J
    from timemachine_b200 import potentials

    s = water_box(700, seed=9)  # box 2.76 nm > 2 cutoff: the energy is a smooth function of the coordinates
    N = s["N"]
    box = s["box"]
    x, params = round_to_f32(s["x"]), round_to_f32(s["params"])
    impl = potentials.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], 2.0, 1.2).to_gpu(precision).unbound_impl
    rng = np.random.default_rng(0)
    dx = rng.normal(size=x.shape)
    dp = rng.normal(size=params.shape) * np.array([1.0, 0.01, 0.05, 0.0])
    dp[params[:, 2] == 0, 2] = 0.0  # eps == 0 switches the LJ term off (k_nonbonded.cuh:232): not differentiable there
    u, t_full = J.unbound_impl_jvp(impl, (x, params, box), (dx, dp, None))
    _, t_x = J.unbound_impl_jvp(impl, (x, params, box), (dx, None, None))
    _, t_p = J.unbound_impl_jvp(impl, (x, params, box), (None, dp, None))
    np.testing.assert_allclose(t_full, t_x + t_p, rtol=1e-12)
    assert u == J.call_unbound_impl(impl, x, params, box)
    # the rule itself, exactly: sum(du_dx * dx) + sum(du_dp * dp) of what execute returns
    du_dx, du_dp, u2 = impl.execute(x, params, box, True, True, True)
    assert u2 == u and t_x == np.sum(du_dx * dx) and t_p == np.sum(du_dp * dp)
    # and it is the directional derivative: central differences through the f64 implementation of the same potential.
    # (Loose: the reference's LJ term is not switched, every O-O pair that crosses the cutoff between x - h dx and x + h dx
    # moves the energy by 8e-4 kJ/mol whatever h is - about 1e-3 of this derivative.)
    ref = potentials.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], 2.0, 1.2).to_gpu(np.float64).unbound_impl
    h = 1e-5
    fd_x = (J.call_unbound_impl(ref, x + h * dx, params, box) - J.call_unbound_impl(ref, x - h * dx, params, box)) / (2 * h)
    fd_p = (J.call_unbound_impl(ref, x, params + h * dp, box) - J.call_unbound_impl(ref, x, params - h * dp, box)) / (2 * h)
    np.testing.assert_allclose(t_x, fd_x, rtol=tol, atol=tol * abs(fd_x))
    np.testing.assert_allclose(t_p, fd_p, rtol=tol, atol=tol * abs(fd_p))
    # bound form
    bound = potentials.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], 2.0, 1.2).to_gpu(precision).bind(params).bound_impl
    ub, tb = J.bound_impl_jvp(bound, (x, box), (dx, None))
    assert ub == u and tb == t_x
    with pytest.raises(RuntimeError, match="box derivatives not supported"):
        J.bound_impl_jvp(bound, (x, box), (dx, np.eye(3)))
