"""GPU: neighbour list parity (reference tests/test_nblist.py): block bounds vs the NumPy restatement, tile membership
in canonical form (per row block, sorted set of column atoms) vs brute force, determinism, row-subset lists, argument
validation messages."""

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import load_reference_ops, water_box

pytestmark = pytest.mark.gpu


def ops():
    from timemachine_b200 import custom_ops

    return custom_ops


def nblist_cls(precision):
    return ops().Neighborlist_f32 if precision == np.float32 else ops().Neighborlist_f64


def test_empty_neighborlist():
    with pytest.raises(RuntimeError, match="Neighborlist N must be at least 1"):
        ops().Neighborlist_f32(0)


@pytest.mark.parametrize("precision,atol,rtol", [(np.float32, 1e-6, 1e-6), (np.float64, 1e-7, 1e-7)])
@pytest.mark.parametrize("size", [12, 128, 156, 298])
def test_block_bounds(precision, atol, rtol, size):
    np.random.seed(2020)
    coords = np.random.randn(size, 3)
    box = np.eye(3) * (np.random.rand(3) + 1)
    ref_ctrs, ref_exts = O.reference_block_bounds(coords, box, 32)
    ctrs, exts = nblist_cls(precision)(size).compute_block_bounds(coords, box, 32)
    np.testing.assert_allclose(ref_ctrs, ctrs, atol=atol, rtol=rtol)
    np.testing.assert_allclose(ref_exts, exts, atol=atol, rtol=rtol)
    with pytest.raises(RuntimeError, match="Block size must be 32."):
        nblist_cls(precision)(size).compute_block_bounds(coords, box, 16)


def canonical(ixn_list):
    return [sorted(set(row)) for row in ixn_list]


def check_membership(got, x, box, cutoff, precision, row_idxs=None, band=5e-6):
    """f64: canonical lists equal to brute force.  f32: equal outside the rounding band - every pair closer than
    cutoff - band is listed and nothing farther than cutoff + band is (membership of a pair within float rounding of the
    cutoff is implementation-defined in f32; forces are unaffected because r_list = cutoff + padding)."""
    got = canonical(got)
    if precision == np.float64:
        ref, margin = O.reference_ixn_list(x, box, cutoff, row_idxs=row_idxs)
        assert margin > 1e-13
        assert got == ref
        return
    strict, _ = O.reference_ixn_list(x, box, cutoff - band, row_idxs=row_idxs)
    loose, _ = O.reference_ixn_list(x, box, cutoff + band, row_idxs=row_idxs)
    assert len(got) == len(strict)
    n_border = 0
    for b, (g, lo, hi) in enumerate(zip(got, strict, loose)):
        assert set(lo) <= set(g) <= set(hi), f"row block {b}"
        n_border += len(set(hi) - set(lo))
    # the band is narrow: only a handful of the ~1e5 listed atoms fall in it
    assert n_border <= max(8, sum(len(h) for h in loose) // 500)


@pytest.mark.parametrize("precision", [np.float32, np.float64])
@pytest.mark.parametrize("n_waters,cutoff", [(11, 0.9), (300, 1.0), (999, 1.2), (999, 0.6)])
def test_neighborlist_membership_bit_exact(precision, n_waters, cutoff):
    sys = water_box(n_waters, seed=n_waters)
    # round to f32 so that both precisions see the same positions (tests/test_nblist.py:109-114 does the same)
    x = sys["x"].astype(np.float32).astype(np.float64)
    box = sys["box"]
    nb = nblist_cls(precision)(len(x))
    got = nb.get_nblist(x, box, cutoff)
    check_membership(got, x, box, cutoff, precision)
    ref, _ = O.reference_ixn_list(x, box, cutoff)
    # no duplicates, determinism (tests/test_nblist.py:258-265)
    for row in got:
        assert len(row) == len(set(row))
    again = nb.get_nblist(x, box, cutoff)
    assert canonical(again) == canonical(got)
    assert nb.get_tile_ixn_count() >= sum((len(r) + 31) // 32 for r in canonical(got)) > 0
    assert nb.get_tile_ixn_count() * 32 <= nb.get_max_ixn_count() + 32 * len(ref)


@pytest.mark.parametrize("precision", [np.float32, np.float64])
def test_neighborlist_wrapped_coordinates(precision, rng):
    """Atoms scattered over several periodic images must give the same canonical list as their home-box images."""
    sys = water_box(400, seed=8)
    x = sys["x"].astype(np.float32).astype(np.float64)
    box = sys["box"]
    L = box[0, 0]
    shift = rng.integers(-3, 4, x.shape) * L
    xs = (x + shift).astype(np.float32).astype(np.float64)
    got = nblist_cls(precision)(len(x)).get_nblist(xs, box, 1.0)
    check_membership(got, xs, box, 1.0, precision, band=3e-5)  # |x| up to ~10 nm: f32 ulp ~1e-6


@pytest.mark.parametrize("precision", [np.float32, np.float64])
def test_neighborlist_row_idxs(precision, rng):
    """Row-subset lists (tests/test_nblist.py:189-234): rows = chosen atoms, columns = the complement."""
    sys = water_box(300, seed=4)
    x = sys["x"].astype(np.float32).astype(np.float64)
    box = sys["box"]
    n = len(x)
    nb = nblist_cls(precision)(n)
    rows = rng.choice(n, 50, replace=False).astype(np.uint32)
    nb.set_row_idxs(rows)
    assert nb.get_num_row_idxs() == 50
    check_membership(nb.get_nblist(x, box, 1.1), x, box, 1.1, precision, row_idxs=rows)
    nb.reset_row_idxs()
    assert nb.get_num_row_idxs() == n
    check_membership(nb.get_nblist(x, box, 1.1), x, box, 1.1, precision)


def test_neighborlist_validation():
    nb = ops().Neighborlist_f32(10)
    with pytest.raises(RuntimeError, match="idxs can't be empty"):
        nb.set_row_idxs(np.array([], dtype=np.uint32))
    with pytest.raises(RuntimeError, match="atom indices must be unique"):
        nb.set_row_idxs(np.array([1, 1], dtype=np.uint32))
    with pytest.raises(RuntimeError, match="number of idxs must be less than N"):
        nb.set_row_idxs(np.arange(10, dtype=np.uint32))
    with pytest.raises(RuntimeError, match="indices values must be less than N"):
        nb.set_row_idxs(np.array([11], dtype=np.uint32))
    with pytest.raises(RuntimeError, match="size is must be at least 1"):
        nb.resize(0)
    with pytest.raises(RuntimeError, match="size is greater than max size: 11 > 10"):
        nb.resize(11)
    with pytest.raises(RuntimeError, match="N != N_"):
        nb.get_nblist(np.zeros((5, 3)), np.eye(3) * 3, 1.0)
    nb.resize(5)
    assert nb.get_nblist(np.zeros((5, 3)) + np.arange(5)[:, None] * 0.1, np.eye(3) * 3, 1.0) == [[0, 1, 2, 3, 4]]


@pytest.mark.parametrize("precision", [np.float32, np.float64])
def test_neighborlist_matches_reference_custom_ops(precision):
    ref = load_reference_ops()
    if ref is None:
        pytest.skip("oracle/_ref/custom_ops*.so not built")
    sys = water_box(999, seed=12)
    x = sys["x"].astype(np.float32).astype(np.float64)
    box = sys["box"]
    suffix = "f32" if precision == np.float32 else "f64"
    ref_list = getattr(ref, f"Neighborlist_{suffix}")(len(x)).get_nblist(x, box, 1.3)
    got = nblist_cls(precision)(len(x)).get_nblist(x, box, 1.3)
    if precision == np.float64:
        assert canonical(got) == canonical(ref_list)
    else:
        check_membership(got, x, box, 1.3, precision)
        check_membership(ref_list, x, box, 1.3, precision)
    rc, re_ = getattr(ref, f"Neighborlist_{suffix}")(len(x)).compute_block_bounds(x, box, 32)
    c, e = nblist_cls(precision)(len(x)).compute_block_bounds(x, box, 32)
    np.testing.assert_allclose(c, rc, rtol=0, atol=1e-6)
    np.testing.assert_allclose(e, re_, rtol=0, atol=1e-6)


def test_tile_buffer_is_sized_from_measured_counts_and_grows(tmp_path):
    """The list buffer starts at 4 tiles per atom instead of the O((N/32)^2) worst case the reference allocates
    (neighborlist.cu:22-28), and a build that needs more grows it and redoes the evaluation: results with a deliberately
    tiny initial buffer (TMB_NBLIST_TILES_PER_ATOM_X100=5: 0.05 tiles per atom) are bitwise those of the default."""
    import subprocess
    import sys

    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from tests.common import water_box, round_to_f32
from timemachine_b200 import custom_ops as ops, potentials as P
s = water_box(4000, seed=3)
N = s["N"]; x, p, box = round_to_f32(s["x"]), round_to_f32(s["params"]), s["box"]
ap = ops.NonbondedAllPairs_f32(N, 2.0, 1.2, None, False, 0.1)
cap0, worst = ap.get_tile_capacity()
dx, dp, u = ap.execute(x, p, box)
cap1, _ = ap.get_tile_capacity()
T = ap.get_tile_count()
nl = ops.Neighborlist_f32(N)
ixn = nl.get_nblist(x, box, 1.3)
# MD through a Context with a list that is too small at first: the call fails loudly once, then runs
pot = P.SummedPotential(
    [P.HarmonicBond(s["bond_idxs"]), P.HarmonicAngle(s["angle_idxs"]), P.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], 2.0, 1.2)],
    [s["bond_params"], s["angle_params"], p])
flat = np.concatenate([s["bond_params"].reshape(-1), s["angle_params"].reshape(-1), p.reshape(-1)])
bp = ops.BoundPotential(pot.to_gpu(np.float32).unbound_impl, flat)
intg = ops.LangevinIntegrator(s["masses"], 300.0, 1e-3, 1.0, 1)
ctx = ops.Context(x, np.zeros_like(x), box, intg, [bp])
msg = ""
try:
    xs, _ = ctx.multiple_steps(20)
except RuntimeError as e:
    msg = str(e)
    ctx.set_x_t(x); ctx.set_v_t(np.zeros_like(x)); intg.set_step(0)  # restore the state, noise stream included
    xs, _ = ctx.multiple_steps(20)
np.savez(sys.argv[1], dx=dx, dp=dp, u=u, cap0=cap0, cap1=cap1, worst=worst, T=T, n_ixn=sum(len(r) for r in ixn), msg=msg, xs=xs)
""" % str(__import__("pathlib").Path(__file__).resolve().parents[1])
    import os

    outs = {}
    for name, env in (("default", {}), ("tiny", {"TMB_NBLIST_TILES_PER_ATOM_X100": "5"})):
        out = tmp_path / f"{name}.npz"
        subprocess.run([sys.executable, "-c", code, str(out)], check=True, env={**os.environ, **env})
        outs[name] = dict(np.load(out))
    d, t = outs["default"], outs["tiny"]
    N = 12000
    assert int(d["cap0"]) == 4 * N + 1024 and int(d["cap1"]) == int(d["cap0"])  # liquid density needs ~1 tile per atom
    assert int(d["worst"]) > int(d["cap0"])  # 1.4x at 12k atoms, 3.6x at 30k, 11x at 90k: the worst case grows as N^2
    assert 0.5 * N < int(d["T"]) < 2 * N and int(d["T"]) * 2 < int(d["cap0"])
    assert int(t["cap0"]) == N * 5 // 100 and int(t["cap1"]) >= int(t["T"]) > int(t["cap0"])  # grown past what was needed
    for k in ("dx", "dp", "u", "T", "n_ixn", "xs"):
        np.testing.assert_array_equal(d[k], t[k], err_msg=k)
    assert str(d["msg"]) == "" and "neighborlist tile buffer overflow during MD steps" in str(t["msg"])
