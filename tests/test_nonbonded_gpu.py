"""GPU parity tests of the nonbonded path, modelled on the reference's tests/nonbonded/*.py:
differential against the CPU oracle for all 8 (du_dx, du_dp, u) combinations, bitwise run-to-run determinism, bitwise
invariance to Hilbert sorting / neighbour-list padding / decomposition into interaction groups, exact cancellation by
exclusions, and - when oracle/_ref holds the compiled reference - forces within 1e-5 of the reference's own kernels."""

import itertools

import numpy as np
import pytest

from tests.common import assert_forces_close, load_reference_ops, random_nonbonded_system, round_to_f32, water_box
from oracle import tm_oracle as O

pytestmark = pytest.mark.gpu

BETA, CUTOFF = 2.0, 1.2


def ops():
    from timemachine_b200 import custom_ops

    return custom_ops


def pots():
    from timemachine_b200 import potentials

    return potentials


def tolerances(precision):
    # f64: the reference's own tolerance (tests/nonbonded/test_nonbonded_all_pairs.py:163).  f32 against the f64 oracle:
    # the reference uses rtol 1e-4 on its test systems; the synthetic systems here (bare charges on eps == 0 atoms at
    # liquid density) have per-atom forces that are sums of ~400 strongly cancelling terms, so f32 round-off
    # (6e-8 * sum|terms|) reaches 2-3e-4 of the NET force.  The tight f32 statement is made against the reference's own
    # f32 kernels (1e-5, test_against_reference_custom_ops), which also checks our f32 error vs f64 is no larger than
    # the reference's.
    return (1e-8, 1e-8) if precision == np.float64 else (5e-4, 5e-4)


def compare_against_oracle(impl, oracle_fn, x, params, box, precision, du_dp_atol_scale=1.0):
    """GradientTest.compare_forces (reference tests/common.py:275-334): all flag combos, twice, bitwise repeatable."""
    rtol, atol = tolerances(precision)
    x, params = round_to_f32(x), round_to_f32(params)  # so f32 and f64 see identical inputs
    ref_u, ref_dx, ref_dp = oracle_fn(x, params, box)
    for want_dx, want_dp, want_u in itertools.product([False, True], repeat=3):
        r1 = impl.execute(x, params, box, want_dx, want_dp, want_u)
        r2 = impl.execute(x, params, box, want_dx, want_dp, want_u)
        for a, b in zip(r1, r2):
            if a is None:
                assert b is None
            else:
                np.testing.assert_array_equal(a, b)
        dx, dp, u = r1
        assert (dx is None) == (not want_dx) and (dp is None) == (not want_dp) and (u is None) == (not want_u)
        if want_u:
            np.testing.assert_allclose(u, ref_u, rtol=rtol, atol=atol)
        if want_dx:
            assert_forces_close(ref_dx, dx, rtol)
        if want_dp:
            assert dp.shape == params.shape
            assert_forces_close(ref_dp, dp, rtol * 10 * du_dp_atol_scale, what="du_dp")


# ---- argument validation: exact messages (tests/nonbonded/test_nonbonded_all_pairs.py:10-50) ------------------------
def test_all_pairs_invalid_atom_idxs():
    P = pots()
    with pytest.raises(RuntimeError, match="indices can't be empty"):
        P.NonbondedAllPairs(3, 2.0, 1.1, []).to_gpu(np.float64)
    with pytest.raises(RuntimeError, match="atom indices must be unique"):
        P.NonbondedAllPairs(3, 2.0, 1.1, [0, 0]).to_gpu(np.float64)
    with pytest.raises(RuntimeError, match="index values must be greater or equal to zero"):
        P.NonbondedAllPairs(3, 2.0, 1.1, [0, -1]).to_gpu(np.float64)
    with pytest.raises(RuntimeError, match="index values must be less than N"):
        P.NonbondedAllPairs(3, 2.0, 1.1, [0, 100]).to_gpu(np.float64)
    impl = P.NonbondedAllPairs(3, 2.0, 1.1).to_gpu(np.float32).unbound_impl
    with pytest.raises(RuntimeError, match="indices can't be empty"):
        impl.set_atom_idxs([])
    with pytest.raises(RuntimeError, match="atom indices must be unique"):
        impl.set_atom_idxs([0, 0])


def test_all_pairs_invalid_sizes():
    impl = pots().NonbondedAllPairs(1, 2.0, 1.1).to_gpu(np.float32).unbound_impl
    with pytest.raises(RuntimeError) as e:
        impl.execute(np.zeros((2, 3)), np.zeros((1, 3)), np.eye(3))
    assert "NonbondedAllPairs::execute_device(): expected N == N_, got N=2, N_=1" == str(e.value)
    with pytest.raises(RuntimeError) as e:
        impl.execute(np.zeros((1, 3)), np.zeros((2, 3)), np.eye(3))
    assert "NonbondedAllPairs::execute_device(): expected P == N_*4, got P=6, N_*4=4" == str(e.value)


def test_all_pairs_get_set_atom_idxs(rng):
    n = 231
    impl = pots().NonbondedAllPairs(n, BETA, 1.1).to_gpu(np.float32).unbound_impl
    assert impl.get_atom_idxs() == list(range(n)) and impl.get_num_atom_idxs() == n
    sub = sorted(rng.choice(n, n // 2, replace=False).tolist())
    impl.set_atom_idxs(sub)
    assert impl.get_atom_idxs() == sub and impl.get_num_atom_idxs() == len(sub)


# ---- correctness vs the oracle -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("n", [33, 65, 231, 1050])
@pytest.mark.parametrize("w_pattern", ["zero", "all_same", "some", "cutoff", "beyond"])
def test_all_pairs_correctness(precision, n, w_pattern):
    if n > 231 and w_pattern not in ("zero", "some"):
        pytest.skip("large sizes cover two w patterns")
    x, params, box = random_nonbonded_system(n, seed=n, w_pattern=w_pattern)
    impl = pots().NonbondedAllPairs(n, BETA, CUTOFF).to_gpu(precision).unbound_impl
    compare_against_oracle(impl, lambda x_, p_, b_: O.nonbonded_all_pairs(x_, p_, b_, BETA, CUTOFF), x, params, box, precision)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_all_pairs_large(precision):
    n = 3080  # the reference's largest JAX-comparable size (tests/nonbonded/test_nonbonded.py:121-122)
    x, params, box = random_nonbonded_system(n, seed=7, w_pattern="some")
    impl = pots().NonbondedAllPairs(n, BETA, CUTOFF).to_gpu(precision).unbound_impl
    rtol, atol = tolerances(precision)
    x, params = round_to_f32(x), round_to_f32(params)
    ref_u, ref_dx, ref_dp = O.nonbonded_all_pairs(x, params, box, BETA, CUTOFF)
    dx, dp, u = impl.execute(x, params, box)
    # f32 vs the f64 oracle: per-atom sums of ~400 partially cancelling terms of relative accuracy ~1e-6 each; the
    # tight f32 bound (1e-5) is the one against the reference's own f32 kernels, test_against_reference_custom_ops
    if precision == np.float32:
        rtol *= 5
    np.testing.assert_allclose(u, ref_u, rtol=rtol, atol=atol * 10)
    assert_forces_close(ref_dx, dx, rtol)
    assert_forces_close(ref_dp, dp, rtol * 10, what="du_dp")


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_all_pairs_atom_subset(precision, rng):
    """Subset via atom_idxs == full evaluation of the extracted subsystem, bitwise
    (tests/nonbonded/test_nonbonded_all_pairs.py:93-158)."""
    n = 300
    x, params, box = random_nonbonded_system(n, seed=11, w_pattern="some")
    sub = np.sort(rng.choice(n, 180, replace=False)).astype(np.int32)
    impl_sub = pots().NonbondedAllPairs(n, BETA, CUTOFF, atom_idxs=sub).to_gpu(precision).unbound_impl
    dx, dp, u = impl_sub.execute(x, params, box)
    impl_small = pots().NonbondedAllPairs(len(sub), BETA, CUTOFF).to_gpu(precision).unbound_impl
    dx2, dp2, u2 = impl_small.execute(x[sub], params[sub], box)
    np.testing.assert_array_equal(dx[sub], dx2)
    np.testing.assert_array_equal(dp[sub], dp2)
    assert u == u2
    rest = np.setdiff1d(np.arange(n), sub)
    assert not dx[rest].any() and not dp[rest].any()
    ref_u, ref_dx, _ = O.nonbonded_all_pairs(x, params, box, BETA, CUTOFF, atom_idxs=sub)
    rtol, atol = tolerances(precision)
    np.testing.assert_allclose(u, ref_u, rtol=rtol * 10, atol=atol)
    assert_forces_close(ref_dx, dx, rtol * 10)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("n", [100, 1500])
def test_hilbert_sort_does_not_change_results(precision, n):
    """tests/nonbonded/test_nonbonded_all_pairs.py:197-233: sorted vs unsorted are bitwise identical."""
    x, params, box = random_nonbonded_system(n, seed=n + 1, w_pattern="some")
    a = pots().NonbondedAllPairs(n, BETA, CUTOFF, disable_hilbert_sort=False).to_gpu(precision).unbound_impl
    b = pots().NonbondedAllPairs(n, BETA, CUTOFF, disable_hilbert_sort=True).to_gpu(precision).unbound_impl
    ra = a.execute(x, params, box)
    rb = b.execute(x, params, box)
    for u, v in zip(ra, rb):
        np.testing.assert_array_equal(u, v)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_rebuild_padding_invariance(precision, rng):
    """tests/nonbonded/test_nonbonded.py:66-117: a padded, lazily rebuilt neighbour list gives bitwise the same forces
    as rebuilding every call (padding = 0) while atoms drift."""
    n = 900
    x, params, box = random_nonbonded_system(n, seed=3)
    lazy = pots().NonbondedAllPairs(n, BETA, CUTOFF, nblist_padding=0.1).to_gpu(precision).unbound_impl
    eager = pots().NonbondedAllPairs(n, BETA, CUTOFF, nblist_padding=0.0).to_gpu(precision).unbound_impl
    for step in range(12):
        r1 = lazy.execute(x, params, box)
        r2 = eager.execute(x, params, box)
        for u, v in zip(r1, r2):
            np.testing.assert_array_equal(u, v)
        x = x + rng.normal(0, 0.012, x.shape)  # several steps stay inside padding/2, then a rebuild triggers
    # a box change must also trigger a rebuild
    box2 = box * 1.02
    for u, v in zip(lazy.execute(x, params, box2), eager.execute(x, params, box2)):
        np.testing.assert_array_equal(u, v)


# ---- exclusions ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_small_box_change_reuses_the_list(precision):
    """A barostat-sized box change (and the matching coordinate scaling) is absorbed by the padding: no rebuild, and the
    result is bitwise the one of a potential that builds its list from scratch at the new box.  A change of the order of
    the padding rebuilds."""
    n = 1500
    x, params, box = random_nonbonded_system(n, seed=21, w_pattern="zero")
    x, params = round_to_f32(x), round_to_f32(params)
    impl = pots().NonbondedAllPairs(n, BETA, CUTOFF).to_gpu(precision).unbound_impl
    impl.execute(x, params, box)
    assert impl.get_num_rebuilds() == 1
    for scale, expect_rebuilds in ((1.002, 1), (0.999, 1), (1.05, 2)):
        x2, box2 = x * scale, box * scale
        got = impl.execute(x2, params, box2)
        assert impl.get_num_rebuilds() == expect_rebuilds, scale
        fresh = pots().NonbondedAllPairs(n, BETA, CUTOFF).to_gpu(precision).unbound_impl.execute(x2, params, box2)
        for a, b in zip(got, fresh):
            np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_nonbonded_with_exclusions_water(precision):
    sys = water_box(120, seed=5)
    N = sys["N"]
    pot = pots().Nonbonded(N, sys["exclusion_idxs"], sys["scale_factors"], BETA, CUTOFF)
    impl = pot.to_gpu(precision).unbound_impl

    def oracle(x_, p_, b_):
        return O.nonbonded(x_, p_, b_, sys["exclusion_idxs"], sys["scale_factors"], BETA, CUTOFF)

    compare_against_oracle(impl, oracle, sys["x"], sys["params"], sys["box"], precision)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_exclusions_cancel_overlapping_atoms_exactly(precision):
    """tests/nonbonded/test_nonbonded.py:208-251 and tests/test_energy_overflows.py:179: a pair at (almost) zero distance
    overflows fixed point in the all-pairs term; with the pair excluded the sum must be finite and exact."""
    n = 64
    x, params, box = random_nonbonded_system(n, seed=21)
    params[:, 2] = np.maximum(params[:, 2], 0.3)
    x[1] = x[0] + 1e-6  # clash
    excl = np.array([[0, 1]], dtype=np.int32)
    scales = np.ones((1, 2))
    all_pairs = pots().NonbondedAllPairs(n, BETA, CUTOFF).to_gpu(precision).unbound_impl
    _, _, u_clash = all_pairs.execute(x, params, box)
    # the clashing term is pinned to LLONG_MAX (k_fixed_point.cuh:88-98): the sum either leaves the int64 range (NaN,
    # wrap_kernels.cpp:83-89) or, when the remaining terms are negative, stays just inside it (~1.3e8 kJ/mol)
    assert np.isnan(u_clash) or u_clash > 1e8
    full = pots().Nonbonded(n, excl, scales, BETA, CUTOFF).to_gpu(precision).unbound_impl
    dx, dp, u = full.execute(x, round_to_f32(params), box)
    assert np.isfinite(u) and np.isfinite(dx).all() and np.isfinite(dp).all()
    # reference value: the same system with the clashing pair simply left out of the sum (a float oracle cannot form
    # inf - inf): pairs among {all atoms but 1}  +  atom 1 against {all atoms but 0 and 1}
    xr, pr = x, round_to_f32(params)
    others = np.array([0] + list(range(2, n)))
    a = O.nonbonded_all_pairs(xr, pr, box, BETA, CUTOFF, atom_idxs=others)
    b = O.nonbonded_interaction_group(xr, pr, box, np.array([1]), np.arange(2, n), BETA, CUTOFF)
    ref_u, ref_dx = a[0] + b[0], a[1] + b[1]
    rtol, atol = tolerances(precision)
    np.testing.assert_allclose(u, ref_u, rtol=rtol, atol=atol)
    assert_forces_close(ref_dx, dx, rtol)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_pair_list_and_negated_pair_list(precision, rng):
    n = 200
    x, params, box = random_nonbonded_system(n, seed=9, w_pattern="some")
    pairs = np.stack([rng.permutation(n)[:150], rng.permutation(n)[:150]], 1).astype(np.int32)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    scales = rng.uniform(0, 1, (len(pairs), 2))
    pos = pots().NonbondedPairList(pairs, scales, BETA, CUTOFF).to_gpu(precision).unbound_impl
    neg = pots().NonbondedExclusions(pairs, scales, BETA, CUTOFF).to_gpu(precision).unbound_impl
    compare_against_oracle(pos, lambda x_, p_, b_: O.nonbonded_pair_list(x_, p_, b_, pairs, scales, BETA, CUTOFF), x, params, box, precision)
    a = pos.execute(round_to_f32(x), round_to_f32(params), box)
    b = neg.execute(round_to_f32(x), round_to_f32(params), box)
    np.testing.assert_array_equal(a[0], -b[0])
    np.testing.assert_array_equal(a[1], -b[1])
    assert a[2] == -b[2]


def test_pair_list_validation():
    o = ops()
    with pytest.raises(RuntimeError, match="pair_idxs.size\\(\\) must be even"):
        o.NonbondedPairList_f32(np.array([0, 1, 2], dtype=np.int32), np.ones((1, 2)), 2.0, 1.2)
    with pytest.raises(RuntimeError, match="illegal pair with src == dst: 3, 3"):
        o.NonbondedPairList_f32(np.array([[3, 3]], dtype=np.int32), np.ones((1, 2)), 2.0, 1.2)
    with pytest.raises(RuntimeError, match="expected same number of pairs and scale tuples"):
        o.NonbondedPairList_f32(np.array([[0, 1]], dtype=np.int32), np.ones((2, 2)), 2.0, 1.2)


# ---- interaction groups ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("n,n_rows", [(120, 7), (700, 60), (700, 350)])
def test_interaction_group_correctness(precision, n, n_rows, rng):
    x, params, box = random_nonbonded_system(n, seed=n + n_rows, w_pattern="some")
    rows = np.sort(rng.choice(n, n_rows, replace=False)).astype(np.int32)
    cols = np.setdiff1d(np.arange(n), rows).astype(np.int32)
    impl = pots().NonbondedInteractionGroup(n, rows, BETA, CUTOFF).to_gpu(precision).unbound_impl
    compare_against_oracle(
        impl, lambda x_, p_, b_: O.nonbonded_interaction_group(x_, p_, b_, rows, cols, BETA, CUTOFF), x, params, box, precision
    )
    # explicit column subset
    cols2 = cols[::2]
    impl2 = pots().NonbondedInteractionGroup(n, rows, BETA, CUTOFF, col_atom_idxs=cols2).to_gpu(precision).unbound_impl
    compare_against_oracle(
        impl2, lambda x_, p_, b_: O.nonbonded_interaction_group(x_, p_, b_, rows, cols2, BETA, CUTOFF), x, params, box, precision
    )


def test_interaction_group_validation():
    P = pots()
    with pytest.raises(RuntimeError, match="row_atom_idxs must be nonempty"):
        P.NonbondedInteractionGroup(3, [], 2.0, 1.1).to_gpu(np.float32)
    with pytest.raises(RuntimeError, match="atom indices must be unique"):
        P.NonbondedInteractionGroup(3, [1, 1], 2.0, 1.1).to_gpu(np.float32)
    with pytest.raises(RuntimeError, match="row and col indices must be disjoint"):
        P.NonbondedInteractionGroup(4, [0, 1], 2.0, 1.1, col_atom_idxs=[1, 2]).to_gpu(np.float32)
    with pytest.raises(RuntimeError, match="col_atom_idxs must be nonempty"):
        P.NonbondedInteractionGroup(3, [0, 1], 2.0, 1.1, col_atom_idxs=[]).to_gpu(np.float32)
    with pytest.raises(RuntimeError, match="must be less then N\\(3\\) row indices"):
        P.NonbondedInteractionGroup(3, [0, 1, 2], 2.0, 1.1, col_atom_idxs=[0]).to_gpu(np.float32)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_decomposition_is_bitwise_consistent(precision, rng):
    """tests/nonbonded/test_consistency.py:26-98: all-pairs over everything equals, BIT FOR BIT,
    AllPairs(host) + AllPairs(ligand) + InteractionGroup(ligand x host): a pair term does not depend on which class,
    which tile or which role (row/column) evaluates it."""
    n, n_lig = 600, 45
    x, params, box = random_nonbonded_system(n, seed=77, w_pattern="some")
    lig = np.sort(rng.choice(n, n_lig, replace=False)).astype(np.int32)
    host = np.setdiff1d(np.arange(n), lig).astype(np.int32)
    P = pots()
    mono = P.NonbondedAllPairs(n, BETA, CUTOFF).to_gpu(precision).unbound_impl
    parts = P.FanoutSummedPotential(
        [
            P.NonbondedAllPairs(n, BETA, CUTOFF, atom_idxs=host),
            P.NonbondedAllPairs(n, BETA, CUTOFF, atom_idxs=lig),
            P.NonbondedInteractionGroup(n, lig, BETA, CUTOFF, col_atom_idxs=host),
        ]
    ).to_gpu(precision).unbound_impl
    ra = mono.execute(x, params, box)
    rb = parts.execute(x, params, box)
    for u, v in zip(ra, rb):
        np.testing.assert_array_equal(u, v)


# ---- against the reference's own kernels (when oracle/_ref was built) ------------------------------------------------------
@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("dist", [0.16, 0.13, 0.11])
def test_large_pair_terms(precision, dist):
    """Pair forces of 1e5..1e8 kJ/mol/nm (close contacts, still inside the 64-bit fixed-point range) exceed the two
    27-bit limbs of the f32 tile kernel's shared accumulators and must take its direct global path (k_nb_tiles_cq.cu)."""
    n = 200
    x, params, box = random_nonbonded_system(n, seed=17, w_pattern="zero")
    for a, b in ((0, 1), (50, 150), (77, 78)):
        x[b] = x[a] + np.array([dist, 0.0, 0.0])
        params[[a, b], 1] = 0.15
        params[[a, b], 2] = 1.0
    x, params = round_to_f32(x), round_to_f32(params)
    impl = pots().NonbondedAllPairs(n, BETA, CUTOFF).to_gpu(precision).unbound_impl
    dx, dp, u = impl.execute(x, params, box)
    ou, odx, odp = O.nonbonded_all_pairs(x, params, box, BETA, CUTOFF)
    assert np.abs(odx).max() > 2.0**17
    rtol = 1e-9 if precision == np.float64 else 5e-4
    assert_forces_close(odx, dx, rtol)
    assert_forces_close(odp, dp, rtol * 10, what="du_dp")
    np.testing.assert_allclose(u, ou, rtol=rtol)


@pytest.mark.parametrize("precision,rtol", [(np.float32, 1e-5), (np.float64, 1e-9)])
@pytest.mark.parametrize("n", [231, 3080])
def test_against_reference_custom_ops(precision, rtol, n):
    ref = load_reference_ops()
    if ref is None:
        pytest.skip("oracle/_ref/custom_ops*.so not built")
    suffix = "f32" if precision == np.float32 else "f64"
    x, params, box = random_nonbonded_system(n, seed=n + 5, w_pattern="some")
    x, params = round_to_f32(x), round_to_f32(params)
    ref_impl = getattr(ref, f"NonbondedAllPairs_{suffix}")(n, BETA, CUTOFF)
    rdx, rdp, ru = ref_impl.execute(x, params, box)
    impl = pots().NonbondedAllPairs(n, BETA, CUTOFF).to_gpu(precision).unbound_impl
    dx, dp, u = impl.execute(x, params, box)
    # north_star: forces within 1e-5 relative of the reference custom_ops (f32)
    assert_forces_close(rdx, dx, rtol)
    assert_forces_close(rdp, dp, rtol * 10, what="du_dp")
    np.testing.assert_allclose(u, ru, rtol=rtol, atol=rtol * 10)
    # and our error against the f64 oracle is no worse than the reference's own
    if n <= 3080:
        _, odx, _ = O.nonbonded_all_pairs(x, params, box, BETA, CUTOFF)
        norms = np.maximum(np.linalg.norm(odx, axis=1), 1.0)
        err_ours = (np.linalg.norm(dx - odx, axis=1) / norms).max()
        err_ref = (np.linalg.norm(rdx - odx, axis=1) / norms).max()
        assert err_ours <= 1.5 * err_ref + 1e-9, (err_ours, err_ref)
    if precision == np.float32:
        # stronger than the stated tolerance: the per-pair rounding sequence is the reference's (nb_math.cuh), every
        # term is rounded to fixed point before it is summed, so forces, du/dp and energy are BIT-identical.  That includes
        # the components a clashing pair of the random system (|force| >= 2^27 kJ/mol/nm) pushes out of the 64-bit
        # fixed-point range: garbage by contract (fixed_point.hpp:11-17), but the reference's float -> int64 conversion
        # is reproduced operation for operation (fixed_point.cuh), so it is the SAME garbage.
        overflowed = (np.abs(rdx) >= 2.0**26).any()
        bad = np.argwhere(dx != rdx)
        detail = [(tuple(ix), int(round(dx[tuple(ix)] * 2**36)), int(round(rdx[tuple(ix)] * 2**36))) for ix in bad[:6]]
        assert len(bad) == 0, f"{len(bad)} of {dx.size} force components differ (index, ours, reference in fixed point): {detail}; overflow present: {overflowed}"
        assert not np.any(dp != rdp), f"{np.count_nonzero(dp != rdp)} of {dp.size} du_dp components differ"
        assert np.array_equal(np.float64(u), np.float64(ru), equal_nan=True)


def test_compaction_queue_kernel_equals_ring_kernel_bitwise(tmp_path):
    """The f32 production kernel (k_nb_tiles_cq.cu) and the ring formulation (k_nb_tiles.cu, selected with
    TMB_NB_RING=1 in a fresh process) evaluate the same pair terms; only the integer summation order differs."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    script = tmp_path / "dump.py"
    script.write_text(
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {str(root)!r})\n"
        "from tests.common import random_nonbonded_system, round_to_f32\n"
        "from timemachine_b200 import potentials as P\n"
        "out = {}\n"
        "for n, pat in ((33, 'zero'), (700, 'some'), (3080, 'some')):\n"
        "    x, p, box = random_nonbonded_system(n, seed=n + 3, w_pattern=pat)\n"
        "    x, p = round_to_f32(x), round_to_f32(p)\n"
        "    impl = P.NonbondedAllPairs(n, 2.0, 1.2).to_gpu(np.float32).unbound_impl\n"
        "    dx, dp, u = impl.execute(x, p, box)\n"
        "    out[f'dx{n}'], out[f'dp{n}'], out[f'u{n}'] = dx, dp, u\n"
        "    rows = np.arange(0, n, 9).astype(np.int32)\n"
        "    g = P.NonbondedInteractionGroup(n, rows, 2.0, 1.2).to_gpu(np.float32).unbound_impl\n"
        "    gdx, gdp, gu = g.execute(x, p, box)\n"
        "    out[f'gdx{n}'], out[f'gdp{n}'], out[f'gu{n}'] = gdx, gdp, gu\n"
        "np.savez(sys.argv[1], **out)\n"
    )
    results = {}
    for mode in ("0", "1"):
        env = dict(os.environ, TMB_NB_RING=mode)
        path = tmp_path / f"out{mode}.npz"
        subprocess.check_call([sys.executable, str(script), str(path)], env=env)
        results[mode] = dict(np.load(path))
    assert set(results["0"]) == set(results["1"])
    for k in results["0"]:
        np.testing.assert_array_equal(results["0"][k], results["1"][k], err_msg=k)
