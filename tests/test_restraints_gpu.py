"""GPU: FlatBottomBond, ChiralAtomRestraint, ChiralBondRestraint, NonbondedPairListPrecomputed
against the CPU oracle on the golden inputs (all flag combinations, bitwise repeatable), against the compiled reference,
and the constructor / execute argument checks with the reference's messages."""

import itertools
from pathlib import Path

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import assert_forces_close, load_reference_ops, round_to_f32

pytestmark = pytest.mark.gpu
G = dict(np.load(Path(__file__).parent / "golden" / "restraints.npz"))
BETA, CUTOFF = float(G["beta"]), float(G["cutoff"])


def ops():
    from timemachine_b200 import custom_ops

    return custom_ops


def cases(o, suffix):
    """name -> (impl, params, oracle(x, params))"""
    x, box = G["x"], G["box"]
    return {
        "flat_bottom": (
            getattr(o, f"FlatBottomBond_{suffix}")(G["fb_idxs"]), G["fb_params"],
            lambda xx, pp: O.flat_bottom_bond(xx, pp, box, G["fb_idxs"]),
        ),
        "chiral_atom": (
            getattr(o, f"ChiralAtomRestraint_{suffix}")(G["quads"]), G["k_atom"],
            lambda xx, pp: O.chiral_atom_restraint(xx, pp, G["quads"]),
        ),
        "chiral_bond": (
            getattr(o, f"ChiralBondRestraint_{suffix}")(G["quads"], G["signs"]), G["k_bond"],
            lambda xx, pp: O.chiral_bond_restraint(xx, pp, G["quads"], G["signs"]),
        ),
        "precomputed": (
            getattr(o, f"NonbondedPairListPrecomputed_{suffix}")(G["pre_idxs"], BETA, CUTOFF), G["pre_params"],
            lambda xx, pp: O.nonbonded_precomputed(xx, pp, box, G["pre_idxs"], BETA, CUTOFF),
        ),
    }


@pytest.mark.parametrize("suffix,rtol", [("f64", 1e-9), ("f32", 2e-4)])
@pytest.mark.parametrize("name", ["flat_bottom", "chiral_atom", "chiral_bond", "precomputed"])
def test_against_oracle(name, suffix, rtol):
    impl, params, oracle = cases(ops(), suffix)[name]
    x, params = round_to_f32(G["x"]), round_to_f32(params)
    ref_u, ref_dx, ref_dp = oracle(x, params)
    for want_dx, want_dp, want_u in itertools.product([False, True], repeat=3):
        r1 = impl.execute(x, params, G["box"], want_dx, want_dp, want_u)
        r2 = impl.execute(x, params, G["box"], want_dx, want_dp, want_u)
        for a, b in zip(r1, r2):
            assert (a is None) == (b is None)
            if a is not None:
                np.testing.assert_array_equal(a, b)
        dx, dp, u = r1
        if want_u:
            np.testing.assert_allclose(u, ref_u, rtol=rtol, atol=rtol)
        if want_dx:
            assert_forces_close(ref_dx, dx, rtol)
        if want_dp:
            assert dp.shape == params.shape
            np.testing.assert_allclose(dp, np.asarray(ref_dp).reshape(params.shape), rtol=rtol * 10, atol=rtol * 10)


@pytest.mark.parametrize("suffix,rtol", [("f32", 1e-5), ("f64", 1e-10)])
@pytest.mark.parametrize("name", ["flat_bottom", "chiral_atom", "chiral_bond", "precomputed"])
def test_against_reference_custom_ops(name, suffix, rtol):
    ref = load_reference_ops()
    if ref is None:
        pytest.skip("oracle/_ref/custom_ops*.so not built")
    impl, params, _ = cases(ops(), suffix)[name]
    rimpl, _, _ = cases(ref, suffix)[name]
    x, params = round_to_f32(G["x"]), round_to_f32(params)
    dx, dp, u = impl.execute(x, params, G["box"])
    rdx, rdp, ru = rimpl.execute(x, params, G["box"])
    if name in ("precomputed", "flat_bottom") and suffix == "f32":
        # the kernel follows the compiled reference's rounded-operation sequence: bitwise
        np.testing.assert_array_equal(dx, rdx)
        np.testing.assert_array_equal(dp, rdp)
        assert u == ru
    np.testing.assert_allclose(u, ru, rtol=rtol, atol=rtol)
    assert_forces_close(rdx, dx, rtol)
    np.testing.assert_allclose(dp, rdp, rtol=rtol * 10, atol=rtol * 10)


def test_precomputed_pair_without_lj_follows_the_compiled_reference():
    """A pair with eps == 0 but q != 0: the compiled reference adds gradients only from inside its Lennard-Jones branch
    (k_nonbonded_precomputed.cuh:150-181), so such a pair has an electrostatic energy but no force and no du/dp.  The
    drop-in returns what the reference returns; TMB_PRECOMPUTED_FULL_GRADIENT=1 (read once per process) keeps the
    gradient the reference's Python potential and the oracle have."""
    import os
    import subprocess
    import sys

    o = ops()
    x = np.array([[0.0, 0.0, 0.0], [0.3, 0.1, 0.0], [0.5, 0.4, 0.1]])
    box = np.eye(3) * 3.0
    pairs = np.array([[0, 1], [1, 2]], dtype=np.int32)
    params = np.array([[1.5, 0.3, 0.0, 0.0], [0.7, 0.15, 0.4, 0.0]])
    ou, odx, odp = O.nonbonded_precomputed(x, params, box, pairs, BETA, CUTOFF)
    ou1, odx1, odp1 = O.nonbonded_precomputed(x, params[1:], box, pairs[1:], BETA, CUTOFF)
    for suffix, rtol in (("f64", 1e-9), ("f32", 2e-4)):
        dx, dp, u = getattr(o, f"NonbondedPairListPrecomputed_{suffix}")(pairs, BETA, CUTOFF).execute(x, params, box)
        np.testing.assert_allclose(u, ou, rtol=rtol)  # both pairs in the energy
        assert_forces_close(odx1, dx, rtol)  # only the pair with LJ in the forces
        np.testing.assert_allclose(dp[1], odp1[0], rtol=rtol * 10)
        assert not dp[0].any()
        ref = load_reference_ops()
        if ref is not None:
            rdx, rdp, ru = getattr(ref, f"NonbondedPairListPrecomputed_{suffix}")(pairs, BETA, CUTOFF).execute(x, params, box)
            if suffix == "f32":
                np.testing.assert_array_equal(dx, rdx)
                np.testing.assert_array_equal(dp, rdp)
                assert u == ru
            else:
                np.testing.assert_allclose(dx, rdx, rtol=1e-10, atol=1e-10)
                np.testing.assert_allclose(u, ru, rtol=1e-10)
    code = (
        "import numpy as np\n"
        "from timemachine_b200 import custom_ops as o\n"
        f"x = np.array({x.tolist()}); box = np.eye(3) * 3.0\n"
        f"pairs = np.array({pairs.tolist()}, dtype=np.int32); params = np.array({params.tolist()})\n"
        f"dx, dp, u = o.NonbondedPairListPrecomputed_f64(pairs, {BETA}, {CUTOFF}).execute(x, params, box)\n"
        "np.save('/tmp/tmb_full_gradient.npy', np.concatenate([dx.ravel(), dp.ravel(), [u]]))\n"
    )
    env = dict(os.environ, TMB_PRECOMPUTED_FULL_GRADIENT="1")
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=str(Path(__file__).resolve().parents[1]))
    got = np.load("/tmp/tmb_full_gradient.npy")
    np.testing.assert_allclose(got[:9].reshape(3, 3), odx, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(got[9:17].reshape(2, 4), odp, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(got[17], ou, rtol=1e-10)
    assert np.linalg.norm(odx[0]) > 1.0  # the force the default drops is not small


def test_argument_checks():
    o = ops()
    with pytest.raises(RuntimeError, match=r"bond_idxs.size\(\) must be exactly 2\*k!"):
        o.FlatBottomBond_f32(np.array([0, 1, 2], dtype=np.int32))
    with pytest.raises(RuntimeError, match="src == dst"):
        o.FlatBottomBond_f32(np.array([[0, 0]], dtype=np.int32))
    with pytest.raises(RuntimeError, match="idxs must be non-negative"):
        o.FlatBottomBond_f32(np.array([[0, -1]], dtype=np.int32))
    with pytest.raises(RuntimeError, match=r"idxs.size\(\) must be exactly 4\*k!"):
        o.ChiralAtomRestraint_f32(np.array([0, 1, 2], dtype=np.int32))
    with pytest.raises(RuntimeError, match=r"idxs.size\(\) must be exactly 4\*R!"):
        o.ChiralBondRestraint_f32(np.array([0, 1, 2], dtype=np.int32), np.array([1], dtype=np.int32))
    with pytest.raises(RuntimeError, match=r"signs.size\(\) must be exactly R!"):
        o.ChiralBondRestraint_f32(np.array([[0, 1, 2, 3]], dtype=np.int32), np.array([1, 1], dtype=np.int32))
    with pytest.raises(RuntimeError, match="signs must be comprised exclusively of 1 or -1"):
        o.ChiralBondRestraint_f32(np.array([[0, 1, 2, 3]], dtype=np.int32), np.array([2], dtype=np.int32))
    with pytest.raises(RuntimeError, match=r"idxs.size\(\) must be exactly 2\*B!"):
        o.NonbondedPairListPrecomputed_f32(np.array([0, 1, 2], dtype=np.int32), BETA, CUTOFF)
    with pytest.raises(RuntimeError, match="illegal pair with src == dst: 3, 3"):
        o.NonbondedPairListPrecomputed_f32(np.array([[3, 3]], dtype=np.int32), BETA, CUTOFF)
    x, box = G["x"], G["box"]
    with pytest.raises(RuntimeError, match=r"FlatBottomBond::execute_device\(\): expected P == 90, got P=3"):
        o.FlatBottomBond_f32(G["fb_idxs"]).execute(x, np.zeros((1, 3)), box)
    with pytest.raises(RuntimeError, match=r"ChiralAtomRestraint::execute_device\(\): expected P == R, got P=2, R=24"):
        o.ChiralAtomRestraint_f32(G["quads"]).execute(x, np.zeros(2), box)
    with pytest.raises(RuntimeError, match=r"ChiralBondRestraint::execute_device\(\): expected P == R, got P=2, R=24"):
        o.ChiralBondRestraint_f32(G["quads"], G["signs"]).execute(x, np.zeros(2), box)
    with pytest.raises(RuntimeError, match=r"expected P == 4\*B, got P=8, 4\*B=240"):
        o.NonbondedPairListPrecomputed_f32(G["pre_idxs"], BETA, CUTOFF).execute(x, np.zeros((2, 4)), box)


def test_dataclass_wrappers_and_summed_potential():
    """The reference-shaped dataclasses construct the same objects and sum with the hot-path potentials."""
    from timemachine_b200 import potentials as P

    x, box = round_to_f32(G["x"]), G["box"]
    pots = [
        P.FlatBottomBond(G["fb_idxs"]), P.ChiralAtomRestraint(G["quads"]), P.ChiralBondRestraint(G["quads"], G["signs"]),
        P.NonbondedPairListPrecomputed(G["pre_idxs"], BETA, CUTOFF),
    ]
    params = [G["fb_params"], G["k_atom"], G["k_bond"], G["pre_params"]]
    total = P.SummedPotential(pots, params).to_gpu(np.float64).unbound_impl
    flat = np.concatenate([np.asarray(p).reshape(-1) for p in params])
    dx, dp, u = total.execute(x, flat, box)
    parts = [p.to_gpu(np.float64).unbound_impl.execute(x, q, box) for p, q in zip(pots, params)]
    np.testing.assert_allclose(u, sum(pt[2] for pt in parts), rtol=1e-12)
    np.testing.assert_array_equal(dx, sum(pt[0] for pt in parts))
    np.testing.assert_array_equal(dp, np.concatenate([pt[1].reshape(-1) for pt in parts]))
