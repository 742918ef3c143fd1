"""GPU: water exchange by biased deletion and the machinery behind it (second half), through the
C ABI, against (a) the NumPy oracle, (b) the compiled, unmodified reference custom_ops (oracle/_ref) and (c) the
invariants the reference's own tests assert (tests/test_cuda_bd_exchange_mover.py, tests/test_exchange_mover.py,
tests/test_cuda_rotations.py, tests/test_segmented_*.py in /root/reference).

Tolerances: integers (fixed-point molecule energies, sampled indices, accepted moves, coordinates after a move with the
same seed) are compared exactly; f32 log weights within 1e-5 relative of the f64 oracle."""

import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import load_reference_ops, water_box

pytestmark = pytest.mark.gpu

TEMP = 300.0
BETA = 2.0
CUTOFF = 1.2


def ops():
    from timemachine_b200 import custom_ops

    return custom_ops


def water_system(n_waters, seed=2023, ions=0):
    """Water box, optionally preceded by `ions` single-atom "solute" atoms that are not target molecules."""
    sys_ = water_box(n_waters, seed=seed)
    x, params, box = sys_["x"], sys_["params"], sys_["box"]
    rng = np.random.default_rng(seed)
    if ions:
        xi = rng.uniform(0, box[0, 0], (ions, 3))
        pi = np.stack([rng.normal(0, 5, ions), rng.uniform(0.1, 0.2, ions), rng.uniform(0.2, 1.0, ions), np.zeros(ions)], 1)
        x, params = np.concatenate([xi, x]), np.concatenate([pi, params])
    mols = [[ions + 3 * i, ions + 3 * i + 1, ions + 3 * i + 2] for i in range(n_waters)]
    return x, params, box, mols


def klass(o, name, precision):
    return getattr(o, f"{name}_{'f32' if precision == np.float32 else 'f64'}")


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_bd_exchange_validation(precision):
    """Error strings of the reference (tests/test_cuda_bd_exchange_mover.py:96-153)."""
    o = ops()
    k = klass(o, "BDExchangeMove", precision)
    N, seed, ppm = 10, 2023, 1
    params = np.random.default_rng(2023).random((N, 4))
    with pytest.raises(RuntimeError, match="must provide at least one molecule"):
        k(N, [], params, TEMP, BETA, CUTOFF, seed, ppm, 1)
    with pytest.raises(RuntimeError, match="Molecules are not contiguous: mol 1"):
        k(N, [[0, 1, 2], [4, 5]], params, TEMP, BETA, CUTOFF, seed, ppm, 1)
    with pytest.raises(RuntimeError, match="only support running with mols with constant size, got mixed sizes"):
        k(N, [[0, 1, 2], [3, 4]], params, TEMP, BETA, CUTOFF, seed, ppm, 1)
    with pytest.raises(RuntimeError, match="must provide non-empty molecule indices"):
        k(N, [[]], params, TEMP, BETA, CUTOFF, seed, ppm, 1)
    with pytest.raises(RuntimeError, match="proposals per move must be greater than 0"):
        k(N, [[]], params, TEMP, BETA, CUTOFF, seed, 0, 1)
    with pytest.raises(RuntimeError, match="must provide interval greater than 0"):
        k(N, [[0], [1]], params, TEMP, BETA, CUTOFF, seed, ppm, 0)
    with pytest.raises(RuntimeError, match="must provide batch size greater than 0"):
        k(N, [[0], [1]], params, TEMP, BETA, CUTOFF, seed, ppm, 1, batch_size=-1)
    with pytest.raises(RuntimeError, match="number of proposals per move must be greater than batch size"):
        k(N, [[0], [1]], params, TEMP, BETA, CUTOFF, seed, ppm, 1, batch_size=ppm + 1)
    with pytest.raises(RuntimeError, match="Number of parameters must match N"):
        k(N + 1, [[0], [1]], params, TEMP, BETA, CUTOFF, seed, ppm, 1)
    mover = k(N, [[0], [1]], params, TEMP, BETA, CUTOFF, seed, ppm, 1)
    # get / set params (reference :158-182)
    np.testing.assert_array_equal(mover.get_params(), params)
    new = params + 1.0
    mover.set_params(new)
    np.testing.assert_array_equal(mover.get_params(), new)
    with pytest.raises(RuntimeError, match="number of params don't match"):
        mover.set_params(new[:5])
    assert mover.n_proposed() == 0 and mover.n_accepted() == 0 and mover.batch_size() == 1
    assert mover.last_log_probability() == 0.0  # log-sum-exp buffers start cleared (bd_exchange_move.cu:89-94)


@pytest.mark.parametrize("precision,rtol", [(np.float64, 1e-9), (np.float32, 2e-5)])
@pytest.mark.parametrize("ions", [0, 7])
def test_mol_energies_and_atom_by_atom_vs_oracle(precision, rtol, ions):
    o = ops()
    x, params, box, mols = water_system(120, ions=ions)
    # a few atoms in other periodic images: results must not care
    x = x.copy()
    x[5::17] += box[0, 0] * np.array([1.0, -2.0, 0.0])
    pot = klass(o, "NonbondedMolEnergyPotential", precision)(len(x), mols, BETA, CUTOFF)
    got = pot.execute(x, params, box)
    ref = O.mol_energies(x, params, box, mols, BETA, CUTOFF)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * 50)
    # sum over molecules counts every water-water pair twice and every water-ion pair once
    targets = np.array(mols[3] + mols[40])
    ab = klass(o, "atom_by_atom_energies", precision)(targets, x, params, box, BETA, CUTOFF)
    assert ab.shape == (6, len(x)) and ab.dtype == precision
    ref_ab = O.pair_energy_matrix(x, params, box, targets, np.arange(len(x)), BETA, CUTOFF)
    off = np.ones_like(ref_ab, dtype=bool)
    off[np.arange(6), targets] = False  # an atom with itself: NaN in both (rsqrt(0)), not compared
    np.testing.assert_allclose(ab[off], ref_ab[off], rtol=rtol * 5, atol=rtol * 5)
    assert np.all(np.isnan(ab[~off]))
    # the molecule energy is the sum of its atoms' rows outside the molecule
    row = np.nansum(np.delete(ab[:3].astype(np.float64), mols[3], axis=1))
    np.testing.assert_allclose(row, got[3], rtol=max(rtol, 1e-6) * 10)


def test_mol_energy_validation():
    o = ops()
    with pytest.raises(RuntimeError, match="must provide at least one target mol"):
        o.NonbondedMolEnergyPotential_f32(10, [], BETA, CUTOFF)
    with pytest.raises(RuntimeError, match="Grouped indices must be between 0 and N"):
        o.NonbondedMolEnergyPotential_f32(10, [[9, 10]], BETA, CUTOFF)
    with pytest.raises(RuntimeError, match="All grouped indices must be unique"):
        o.NonbondedMolEnergyPotential_f32(10, [[1, 2], [2, 3]], BETA, CUTOFF)
    pot = o.NonbondedMolEnergyPotential_f32(10, [[0, 1]], BETA, CUTOFF)
    with pytest.raises(RuntimeError, match="params N != coords N"):
        pot.execute(np.zeros((10, 3)), np.zeros((9, 4)), np.eye(3))


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_against_reference_building_blocks(precision):
    """Same inputs -> same numbers as the compiled reference: molecule energies (fixed point: exact), atom-by-atom
    energies, rotations, rotate-and-translate (bitwise: the arithmetic is transcribed from the reference's SASS in f32)."""
    ref = load_reference_ops()
    if ref is None:
        pytest.skip("compiled reference not present (oracle/_ref)")
    o = ops()
    exact = precision == np.float32
    x, params, box, mols = water_system(150, ions=5)
    a = klass(o, "NonbondedMolEnergyPotential", precision)(len(x), mols, BETA, CUTOFF).execute(x, params, box)
    b = klass(ref, "NonbondedMolEnergyPotential", precision)(len(x), mols, BETA, CUTOFF).execute(x, params, box)
    if exact:
        np.testing.assert_array_equal(a, b)
    else:
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-9)
    targets = np.array(mols[0] + mols[77], dtype=np.int32)
    a = klass(o, "atom_by_atom_energies", precision)(targets, x, params, box, BETA, CUTOFF)
    b = klass(ref, "atom_by_atom_energies", precision)(targets, x, params, box, BETA, CUTOFF)
    if exact:
        np.testing.assert_array_equal(a, b)
    else:
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-9, equal_nan=True)
    rng = np.random.default_rng(5)
    quats = rng.normal(size=(40, 4))
    coords = rng.normal(size=(9, 3)) * 3.0
    a = klass(o, "rotate_coords", precision)(coords, quats)
    b = klass(ref, "rotate_coords", precision)(coords, quats)
    assert a.shape == b.shape == (9, 40, 3)
    if exact:
        np.testing.assert_array_equal(a, b)
    else:
        np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-13)
    trans = rng.uniform(-1.5, 2.5, size=(40, 3))
    mol = x[mols[11]]
    a = klass(o, "rotate_and_translate_mol", precision)(mol, box, quats, trans)
    b = klass(ref, "rotate_and_translate_mol", precision)(mol, box, quats, trans)
    assert a.shape == b.shape == (40, 3, 3)
    if exact:
        np.testing.assert_array_equal(a, b)
    else:
        np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("precision,atol", [(np.float64, 1e-12), (np.float32, 2e-6)])
def test_rotations_vs_oracle(precision, atol):
    """tests/test_cuda_rotations.py of the reference: rotation by a quaternion, identity, centroid placement."""
    o = ops()
    rng = np.random.default_rng(11)
    coords = rng.normal(size=(6, 3))
    quats = rng.normal(size=(25, 4))
    got = klass(o, "rotate_coords", precision)(coords, quats)
    for r, q in enumerate(quats):
        np.testing.assert_allclose(got[:, r], O.quaternion_rotate(coords, q), atol=atol * 10)
    ident = klass(o, "rotate_coords", precision)(coords, np.array([[1.0, 0, 0, 0], [-3.0, 0, 0, 0]]))
    np.testing.assert_allclose(ident[:, 0], coords, atol=atol)
    np.testing.assert_allclose(ident[:, 1], coords, atol=atol)
    box = np.diag([3.0, 4.0, 5.0])
    trans = rng.uniform(-1.0, 2.0, size=(25, 3))
    moved = klass(o, "rotate_and_translate_mol", precision)(coords, box, quats, trans)
    for b in range(25):
        np.testing.assert_allclose(moved[b], O.rotate_and_translate_mol(coords, box, quats[b], trans[b]), atol=atol * 20)
        c = moved[b].mean(0)
        assert np.all(c > -1e-5) and np.all(c < np.diag(box) + 1e-5)
    with pytest.raises(RuntimeError, match="quaternions must have a shape that is 4 dimensional"):
        klass(o, "rotate_coords", precision)(coords, np.zeros((3, 3)))
    with pytest.raises(RuntimeError, match="Number of quaternions and translations must match"):
        klass(o, "rotate_and_translate_mol", precision)(coords, box, quats, trans[:3])


@pytest.mark.parametrize("precision,rtol", [(np.float64, 1e-12), (np.float32, 2e-6)])
def test_segmented_sumexp(precision, rtol):
    """tests/test_segmented_sumexp.py of the reference: ragged segments, large magnitudes, error strings."""
    o = ops()
    rng = np.random.default_rng(3)
    k = klass(o, "SegmentedSumExp", precision)
    s = k(2000, 6)
    vals = [rng.normal(0, 30, n) for n in (1, 7, 256, 257, 1999, 2000)]
    got = s.logsumexp(vals)
    ref = [O.logsumexp(np.asarray(v, dtype=precision)) for v in vals]
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol)
    big = s.logsumexp([np.array([1e4, 1e4 - 1.0]), np.array([-1e4, -1e4])])
    np.testing.assert_allclose(big, [1e4 + np.log1p(np.exp(-1.0)), -1e4 + np.log(2.0)], rtol=rtol)
    with pytest.raises(RuntimeError, match="empty array not allowed"):
        s.logsumexp([[1.0], []])
    with pytest.raises(RuntimeError, match="number of segments must be less than or equal"):
        s.logsumexp([[1.0]] * 7)
    with pytest.raises(RuntimeError, match="total values is greater than buffer size"):
        k(2, 2).logsumexp([[1.0, 2.0, 3.0], [1.0, 2.0]])


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_segmented_weighted_random_sampler(precision):
    """tests/test_segmented_weighted_random_sampler.py of the reference: validation, determinism per seed, zero-weight
    entries never drawn, empirical frequencies follow the weights; and the same draws as the compiled reference."""
    o = ops()
    k = klass(o, "SegmentedWeightedRandomSampler", precision)
    rng = np.random.default_rng(9)
    n, segs = 50, 400
    w = rng.random(n)
    w[::5] = 0.0
    weights = [w] * segs
    a = k(n, segs, 2024).sample(weights)
    b = k(n, segs, 2024).sample(weights)
    c = k(n, segs, 2025).sample(weights)
    assert a == b and a != c
    assert all(0 <= v < n and w[v] > 0 for v in a)
    sampler = k(n, segs, 1)
    draws = np.concatenate([sampler.sample(weights) for _ in range(50)])
    freq = np.bincount(draws, minlength=n) / len(draws)
    np.testing.assert_allclose(freq, w / w.sum(), atol=0.006)
    ragged = [rng.random(m) for m in rng.integers(1, n + 1, segs)]
    r = k(n, segs, 7).sample(ragged)
    assert all(0 <= v < len(seg) for v, seg in zip(r, ragged))
    for bad, msg in (([np.inf, 1.0], "unable to use infinity as a weight"), ([np.nan, 1.0], "unable to use nan as a weight"),
                     ([-1.0, 1.0], "unable to use negative values as a weight"), ([], "empty probability distribution not allowed")):
        with pytest.raises(RuntimeError, match=msg):
            k(2, 1, 1).sample([bad])
    with pytest.raises(RuntimeError, match="number of segments don't match"):
        k(n, segs, 1).sample(weights[:3])
    ref = load_reference_ops()
    if ref is not None:
        rs = klass(ref, "SegmentedWeightedRandomSampler", precision)(n, segs, 2024)
        mine = k(n, segs, 2024)
        for _ in range(3):  # successive calls continue the same cuRAND stream
            assert mine.sample(weights) == list(rs.sample(weights))


@pytest.mark.parametrize("precision,rtol", [(np.float64, 1e-9), (np.float32, 2e-5)])
@pytest.mark.parametrize("batch_size", [1, 5])
def test_log_weights_initial_and_incremental(precision, rtol, batch_size):
    """compute_initial_log_weights against the oracle; compute_incremental_log_weights equals the initial weights
    recomputed on the moved coordinates (the reference's own property test, test_cuda_bd_exchange_mover.py:607-669) and
    the oracle; against the compiled reference the f32 weights agree exactly."""
    o = ops()
    x, params, box, mols = water_system(90, ions=4)
    N = len(x)
    k = klass(o, "BDExchangeMove", precision)
    mover = k(N, mols, params, TEMP, BETA, CUTOFF, 2023, batch_size, 1, batch_size=batch_size)
    w0 = np.array(mover.compute_initial_log_weights(x, box), dtype=np.float64)
    ref0 = O.bd_log_weights(x, params, box, mols, BETA, CUTOFF, TEMP)
    np.testing.assert_allclose(w0, ref0, rtol=rtol, atol=rtol * 100)
    np.testing.assert_array_equal(mover.get_before_log_weights(), w0.astype(precision))

    rng = np.random.default_rng(4)
    idxs = rng.choice(len(mols), batch_size, replace=False).astype(np.int32)
    quats = rng.normal(size=(batch_size, 4))
    trans = rng.uniform(0, 1, (batch_size, 3)) * np.diag(box)  # used as given (not scaled)
    inc = mover.compute_incremental_log_weights(x, box, idxs, quats, trans)
    assert len(inc) == batch_size and all(len(r) == len(mols) for r in inc)
    rotate = klass(o, "rotate_and_translate_mol", precision)
    for b in range(batch_size):
        moved = x.copy()
        moved[mols[idxs[b]]] = rotate(x[mols[idxs[b]]], box, quats[b : b + 1], trans[b : b + 1] / np.diag(box))[0]
        again = np.array(mover.compute_initial_log_weights(moved, box), dtype=np.float64)
        got = np.array(inc[b], dtype=np.float64)
        if precision == np.float64:
            np.testing.assert_allclose(got, again, rtol=1e-10, atol=1e-8)
        else:
            # integer energies: incremental and recomputed weights differ only where translation / box rounds differently
            np.testing.assert_allclose(got, again, rtol=2e-4, atol=2e-3)
        want = O.bd_log_weights(moved, params, box, mols, BETA, CUTOFF, TEMP)
        sane = np.abs(want) < 1e6  # a random placement can clash: pair terms beyond 2^27 kJ/mol saturate in fixed point by design
        assert sane.sum() >= len(mols) - 4
        np.testing.assert_allclose(got[sane], want[sane], rtol=max(rtol, 2e-4), atol=2e-3)
    ref = load_reference_ops()
    if ref is not None:
        rm = klass(ref, "BDExchangeMove", precision)(N, mols, params, TEMP, BETA, CUTOFF, 2023, batch_size, 1, batch_size=batch_size)
        r0 = np.array(rm.compute_initial_log_weights(x, box))
        rinc = np.array(rm.compute_incremental_log_weights(x, box, idxs, quats, trans))
        if precision == np.float32:
            np.testing.assert_array_equal(w0.astype(np.float32), r0)
            np.testing.assert_array_equal(np.array(inc, dtype=np.float32), rinc)
        else:
            np.testing.assert_allclose(w0, r0, rtol=1e-11, atol=1e-9)
            np.testing.assert_allclose(np.array(inc), rinc, rtol=1e-11, atol=1e-8)


@pytest.mark.parametrize("proposals_per_move,batch_size", [(1, 1), (10, 1), (2, 2), (100, 100), (300, 77)])
@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_bd_exchange_deterministic_moves(proposals_per_move, batch_size, precision):
    """The reference's determinism contract (test_cuda_bd_exchange_mover.py:367-424): K moves of one proposal and one
    move of K proposals in batches produce the SAME coordinates and counters, bit for bit."""
    o = ops()
    x, params, _, mols = water_system(60)
    box = np.eye(3) * 100.0  # vacuum: most proposals are accepted
    N = len(x)
    k = klass(o, "BDExchangeMove", precision)
    a = k(N, mols, params, TEMP, BETA, CUTOFF, 2023, 1, 1)
    b = k(N, mols, params, TEMP, BETA, CUTOFF, 2023, proposals_per_move, 1, batch_size=batch_size)
    xa = x.copy()
    for _ in range(proposals_per_move):
        xa, box_a = a.move(xa, box)
        np.testing.assert_array_equal(box_a, box)
    xb, _ = b.move(x, box)
    assert a.n_accepted() >= max(proposals_per_move // 4, 1)
    assert not np.all(xa == x)
    assert a.n_proposed() == b.n_proposed() == proposals_per_move
    assert a.n_accepted() == b.n_accepted()
    np.testing.assert_array_equal(xa, xb)
    # only whole molecules moved, rigidly
    changed = np.any(xa != x, axis=1).reshape(-1, 3)
    assert np.all(changed.all(axis=1) | (~changed).all(axis=1))
    for m in np.nonzero(changed[:, 0])[0]:
        np.testing.assert_allclose(np.linalg.norm(xa[mols[m][0]] - xa[mols[m][1]]), 0.09572, atol=2e-4)  # f32 positions in a 100 nm box resolve 4e-6 nm


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_bd_exchange_repeated_batches_equal_one_long_move(precision):
    """test_cuda_bd_exchange_mover.py:427-487: the same batch size, `iterations` calls of P proposals vs one call of
    iterations * P proposals."""
    o = ops()
    x, params, _, mols = water_system(60)
    box = np.eye(3) * 100.0
    N, P, B, iterations = len(x), 64, 16, 3
    k = klass(o, "BDExchangeMove", precision)
    a = k(N, mols, params, TEMP, BETA, CUTOFF, 2024, P, 1, batch_size=B)
    b = k(N, mols, params, TEMP, BETA, CUTOFF, 2024, P * iterations, 1, batch_size=B)
    xa = x.copy()
    for _ in range(iterations):
        xa, _ = a.move(xa, box)
    xb, _ = b.move(x, box)
    assert a.n_accepted() == b.n_accepted() > 0
    assert a.n_proposed() == b.n_proposed() == P * iterations
    np.testing.assert_array_equal(xa, xb)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("proposals_per_move,batch_size,ions", [(1, 1, 0), (40, 1, 3), (200, 50, 3), (1000, 250, 0)])
def test_bd_moves_against_compiled_reference(precision, proposals_per_move, batch_size, ions):
    """Same seed, same system: the moves of the compiled reference, in bulk water where most proposals are rejected.
    f32 (the precision the reference runs in production): identical coordinates, counters and last log probability up to
    the rounding of the log-sum-exp (the molecule energies are the same integers)."""
    ref = load_reference_ops()
    if ref is None:
        pytest.skip("compiled reference not present (oracle/_ref)")
    o = ops()
    x, params, box, mols = water_system(200, ions=ions)
    N = len(x)
    args = (N, mols, params, TEMP, BETA, CUTOFF, 2023, proposals_per_move, 1)
    mine = klass(o, "BDExchangeMove", precision)(*args, batch_size=batch_size)
    theirs = klass(ref, "BDExchangeMove", precision)(*args, batch_size=batch_size)
    xa, xb = x.copy(), x.copy()
    n_moves = 6 if proposals_per_move < 100 else 3
    for _ in range(n_moves):
        xa, _ = mine.move(xa, box)
        xb, _ = theirs.move(xb, box)
        assert mine.n_accepted() == theirs.n_accepted()
        if precision == np.float32:
            np.testing.assert_array_equal(xa, xb)
        else:
            np.testing.assert_allclose(xa, xb, rtol=0, atol=1e-12)
        np.testing.assert_allclose(mine.last_raw_log_probability(), theirs.last_raw_log_probability(), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(mine.last_log_probability(), theirs.last_log_probability(), rtol=1e-4, atol=1e-4)
    assert mine.n_proposed() == theirs.n_proposed() == n_moves * proposals_per_move
    np.testing.assert_allclose(mine.get_before_log_weights(), theirs.get_before_log_weights(), rtol=1e-6)


def test_bd_move_log_probability_matches_oracle():
    """One proposal per move: the reported acceptance probability is min(lse(before) - lse(after), 0) of the oracle's
    weights, where `after` is evaluated on the proposed coordinates (exchange_mover.py:203-234)."""
    o = ops()
    x, params, box, mols = water_system(80, ions=2)
    N = len(x)
    mover = o.BDExchangeMove_f64(N, mols, params, TEMP, BETA, CUTOFF, 77, 1, 1)
    for _ in range(4):
        before = O.bd_log_weights(x, params, box, mols, BETA, CUTOFF, TEMP)
        x_new, _ = mover.move(x, box)
        raw = mover.last_raw_log_probability()
        after_w = np.array(mover.get_after_log_weights())
        np.testing.assert_allclose(raw, O.logsumexp(before) - O.logsumexp(after_w), rtol=1e-9, atol=1e-9)
        assert mover.last_log_probability() == min(raw, 0.0)
        if mover.n_accepted() and np.any(x_new != x):
            # accepted: the after weights are the weights of the new coordinates
            np.testing.assert_allclose(after_w, O.bd_log_weights(x_new, params, box, mols, BETA, CUTOFF, TEMP), rtol=1e-8, atol=1e-6)
        x = x_new


def test_device_loop_equals_host_driven_loop(tmp_path):
    """The persistent cooperative kernel (default) and the one-launch-per-phase host loop (TMB_BD_LOOP=host) run the same
    phase functions: identical coordinates and counters."""
    script = tmp_path / "run.py"
    script.write_text(
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from timemachine_b200 import custom_ops as o\n"
        "from tests.test_exchange_gpu import water_system, TEMP, BETA, CUTOFF\n"
        "x, params, box, mols = water_system(150, ions=3)\n"
        "out = {}\n"
        "for name, k in (('f32', o.BDExchangeMove_f32), ('f64', o.BDExchangeMove_f64)):\n"
        "    m = k(len(x), mols, params, TEMP, BETA, CUTOFF, 5, 300, 1, batch_size=64)\n"
        "    xs = x\n"
        "    for _ in range(3):\n"
        "        xs, _ = m.move(xs, box)\n"
        "    out[name] = xs; out[name + '_acc'] = m.n_accepted(); out[name + '_lp'] = m.last_raw_log_probability()\n"
        "np.savez(sys.argv[1], **out)\n" % str(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    )
    res = {}
    for mode in ("device", "host"):
        env = dict(os.environ, TMB_BD_LOOP=mode)
        path = tmp_path / f"{mode}.npz"
        subprocess.check_call([sys.executable, str(script), str(path)], env=env, timeout=600)
        res[mode] = dict(np.load(path))
    for key in res["device"]:
        np.testing.assert_array_equal(res["device"][key], res["host"][key], err_msg=key)
    assert res["device"]["f32_acc"] > 0


def test_bd_exchange_mover_in_context():
    """As a Context mover (test_cuda_bd_exchange_mover.py:296-364): MD with an exchange move every few steps; the moved
    waters trigger neighbour-list rebuilds and the trajectory stays finite with rigid-ish waters."""
    o = ops()
    from timemachine_b200 import lib as tmlib
    from timemachine_b200 import potentials as P

    sys_ = water_box(900, seed=7)
    x, box, params, N = sys_["x"], sys_["box"], sys_["params"], sys_["N"]
    mols = [[3 * i, 3 * i + 1, 3 * i + 2] for i in range(900)]
    bps = [
        P.HarmonicBond(sys_["bond_idxs"]).bind(sys_["bond_params"]).to_gpu(np.float32).bound_impl,
        P.HarmonicAngle(sys_["angle_idxs"]).bind(sys_["angle_params"]).to_gpu(np.float32).bound_impl,
        P.Nonbonded(N, sys_["exclusion_idxs"], sys_["scale_factors"], BETA, CUTOFF).bind(params).to_gpu(np.float32).bound_impl,
    ]
    intg = tmlib.LangevinIntegrator(TEMP, 1.0e-3, 1.0, sys_["masses"], 2023).impl()
    mover = o.BDExchangeMove_f32(N, mols, params, TEMP, BETA, CUTOFF, 2023, 50, 5, batch_size=10)
    ctxt = o.Context(x, np.zeros_like(x), box, intg, bps, movers=[mover])
    xs, boxes = ctxt.multiple_steps(100, 10)
    assert xs.shape == (10, N, 3) and np.all(np.isfinite(xs))
    assert mover.n_proposed() == 20 * 50
    assert 0 < mover.n_accepted() < mover.n_proposed()
    d = np.linalg.norm(xs[-1][0::3] - xs[-1][1::3], axis=1)
    assert np.all(np.abs(d - 0.09572) < 0.03)
    # some water jumped further than thermal motion allows in 100 fs
    jump = np.linalg.norm(xs[-1][0::3] - x[0::3], axis=1)
    assert jump.max() > 0.5


# ---------------------------------------------------------------------------------------------------------------------
# targeted insertion / biased deletion (reference tests/test_cuda_targeted_insertion_mover.py)
def ligand_water_system(n_waters, n_ligand_waters=4, seed=2023):
    """A water box whose `n_ligand_waters` molecules nearest to the box centre play the ligand (no clashes, so no
    saturated energies, which the reference's device asserts reject); the ligand atoms come first."""
    s = water_box(n_waters + n_ligand_waters, seed=seed)
    x, params, box = s["x"], s["params"], s["box"]
    c = x.reshape(-1, 3, 3).mean(1)
    order = np.argsort(np.linalg.norm(c - np.diag(box) * 0.5, axis=1), kind="stable")
    atoms = (order[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
    x, params = x[atoms], params[atoms]
    n_lig = 3 * n_ligand_waters
    mols = [[n_lig + 3 * i, n_lig + 3 * i + 1, n_lig + 3 * i + 2] for i in range(n_waters)]
    return x, params, box, mols, np.arange(n_lig, dtype=np.int32)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("radius", [0.4, 0.9, 2.0])
def test_inner_and_outer_mols(precision, radius):
    """test_cuda_targeted_insertion_mover.py:190-219: against the Python reference's get_water_groups."""
    o = ops()
    x, params, box, mols, lig = ligand_water_system(300)
    x = x.copy()
    x[3 * 50 + 12 :: 41] += box[0, 0]  # atoms in another image
    inner, outer = klass(o, "inner_and_outer_mols", precision)(lig, x, box, mols, radius)
    center = x[lig].mean(0)
    ref_in, ref_out = O.water_groups(x, box, center, mols, radius)
    # molecules within float rounding of the sphere may land on either side
    c = np.array([x[m].mean(0) for m in mols])
    d = np.linalg.norm(O.delta_r(c, center, box), axis=1)
    sure = np.abs(d - radius) > 1e-5
    assert set(np.array(inner)[sure[inner]]) == set(ref_in[sure[ref_in]])
    assert set(np.array(outer)[sure[outer]]) == set(ref_out[sure[ref_out]])
    assert sorted(inner + outer) == list(range(len(mols)))
    ref = load_reference_ops()
    if ref is not None:
        r_in, r_out = klass(ref, "inner_and_outer_mols", precision)(lig, x, box, mols, radius)
        assert list(r_in) == inner and list(r_out) == outer


@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("n_translations", [1, 33, 1000])
def test_translations_inside_and_outside_sphere(precision, n_translations):
    """test_cuda_targeted_insertion_mover.py:223-249: inner translations lie in the sphere, outer ones outside it (under
    PBC); deterministic per seed; the same numbers as the compiled reference."""
    o = ops()
    box = np.diag([3.0, 3.5, 4.0])
    center = np.array([1.4, 1.9, 2.2])
    radius = 0.8
    f = klass(o, "translations_inside_and_outside_sphere_host", precision)
    t = f(n_translations, box, center, radius, 2023)
    assert t.shape == (n_translations, 2, 3) and t.dtype == precision
    d_in = np.linalg.norm(O.delta_r(t[:, 0].astype(np.float64), center, box), axis=1)
    d_out = np.linalg.norm(O.delta_r(t[:, 1].astype(np.float64), center, box), axis=1)
    assert np.all(d_in < radius + 1e-5) and np.all(d_out >= radius - 1e-5)
    np.testing.assert_array_equal(t, f(n_translations, box, center, radius, 2023))
    assert not np.array_equal(t, f(n_translations, box, center, radius, 2024))
    if n_translations == 1000:  # uniform in the sphere: the radius^3 law
        np.testing.assert_allclose(np.mean((d_in / radius) ** 3), 0.5, atol=0.04)
    with pytest.raises(RuntimeError, match="Center must be of length 3"):
        f(3, box, np.zeros(2), radius, 1)
    ref = load_reference_ops()
    if ref is not None:
        r = klass(ref, "translations_inside_and_outside_sphere_host", precision)(n_translations, box, center, radius, 2023)
        if precision == np.float32:
            np.testing.assert_array_equal(t, r)
        else:
            np.testing.assert_allclose(t, r, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("precision", [np.float64, np.float32])
def test_tibd_exchange_validation(precision):
    """test_cuda_targeted_insertion_mover.py:253-340."""
    o = ops()
    k = klass(o, "TIBDExchangeMove", precision)
    N, seed, ppm, radius = 10, 2023, 1, 1.0
    params = np.random.default_rng(2023).random((N, 4))
    lig = [0]
    with pytest.raises(RuntimeError, match="must provide at least one molecule"):
        k(N, lig, [], params, TEMP, BETA, CUTOFF, radius, seed, ppm, 1)
    with pytest.raises(RuntimeError, match="must provide at least one atom for the ligand indices"):
        k(N, [], [[1], [2]], params, TEMP, BETA, CUTOFF, radius, seed, ppm, 1)
    with pytest.raises(RuntimeError, match="Molecules are not contiguous: mol 1"):
        k(N, lig, [[1, 2, 3], [5, 6]], params, TEMP, BETA, CUTOFF, radius, seed, ppm, 1)
    with pytest.raises(RuntimeError, match="only support running with mols with constant size, got mixed sizes"):
        k(N, lig, [[1, 2, 3], [4, 5]], params, TEMP, BETA, CUTOFF, radius, seed, ppm, 1)
    with pytest.raises(RuntimeError, match="must provide non-empty molecule indices"):
        k(N, lig, [[]], params, TEMP, BETA, CUTOFF, radius, seed, ppm, 1)
    with pytest.raises(RuntimeError, match="proposals per move must be greater than 0"):
        k(N, lig, [[1], [2]], params, TEMP, BETA, CUTOFF, radius, seed, 0, 1)
    with pytest.raises(RuntimeError, match="radius must be greater than 0.0"):
        k(N, lig, [[1], [2]], params, TEMP, BETA, CUTOFF, 0.0, seed, ppm, 1)
    with pytest.raises(RuntimeError, match="must provide interval greater than 0"):
        k(N, lig, [[1], [2]], params, TEMP, BETA, CUTOFF, radius, seed, ppm, 0)
    with pytest.raises(RuntimeError, match="must provide batch size greater than 0"):
        k(N, lig, [[1], [2]], params, TEMP, BETA, CUTOFF, radius, seed, ppm, 1, batch_size=0)
    with pytest.raises(RuntimeError, match="number of proposals per move must be greater than batch size"):
        k(N, lig, [[1], [2]], params, TEMP, BETA, CUTOFF, radius, seed, ppm, 1, batch_size=2)
    mover = k(N, lig, [[1], [2]], params, TEMP, BETA, CUTOFF, radius, seed, ppm, 1)
    assert mover.last_log_probability() == 0.0 and mover.n_proposed() == 0
    with pytest.raises(RuntimeError, match="volume of inner radius greater than box volume"):
        mover.move(np.random.default_rng(1).random((N, 3)), np.eye(3) * 0.5)
    np.testing.assert_array_equal(mover.get_params(), params)
    with pytest.raises(RuntimeError, match="number of params don't match"):
        mover.set_params(params[:3])


@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("radius,proposals_per_move,batch_size", [(0.5, 1, 1), (0.9, 50, 1), (0.9, 200, 64), (1.4, 120, 120)])
def test_tibd_exchange_deterministic_moves(precision, radius, proposals_per_move, batch_size):
    """test_cuda_targeted_insertion_mover.py:737-822: K moves of one proposal == one move of K proposals in batches,
    bit for bit; every accepted move switches a molecule between the sphere and the bulk."""
    o = ops()
    x, params, box, mols, lig = ligand_water_system(600)
    N = len(x)
    k = klass(o, "TIBDExchangeMove", precision)
    a = k(N, lig, mols, params, TEMP, BETA, CUTOFF, radius, 2023, 1, 1)
    b = k(N, lig, mols, params, TEMP, BETA, CUTOFF, radius, 2023, proposals_per_move, 1, batch_size=batch_size)
    inner_of = klass(o, "inner_and_outer_mols", precision)
    xa = x.copy()
    for _ in range(proposals_per_move):
        before = set(inner_of(lig, xa, box, mols, radius)[0])
        acc = a.n_accepted()
        xa, _ = a.move(xa, box)
        after = set(inner_of(lig, xa, box, mols, radius)[0])
        if a.n_accepted() > acc:
            assert len(before ^ after) == 1  # exactly one molecule changed region
        else:
            assert before == after
    xb, _ = b.move(x, box)
    assert a.n_proposed() == b.n_proposed() == proposals_per_move
    assert a.n_accepted() == b.n_accepted()
    np.testing.assert_array_equal(xa, xb)
    np.testing.assert_array_equal(xa[: len(lig)], x[: len(lig)])  # the ligand never moves


@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("radius,proposals_per_move,batch_size", [(0.6, 1, 1), (0.9, 100, 1), (0.9, 400, 100), (1.5, 1000, 250)])
def test_tibd_moves_against_compiled_reference(precision, radius, proposals_per_move, batch_size):
    """Same seed, same system: the targeted moves of the compiled reference (f32: identical coordinates and counters)."""
    ref = load_reference_ops()
    if ref is None:
        pytest.skip("compiled reference not present (oracle/_ref)")
    o = ops()
    x, params, box, mols, lig = ligand_water_system(700)
    N = len(x)
    args = (N, lig, mols, params, TEMP, BETA, CUTOFF, radius, 2023, proposals_per_move, 1)
    mine = klass(o, "TIBDExchangeMove", precision)(*args, batch_size=batch_size)
    theirs = klass(ref, "TIBDExchangeMove", precision)(*args, batch_size=batch_size)
    xa, xb = x.copy(), x.copy()
    for _ in range(6 if proposals_per_move < 200 else 3):
        xa, _ = mine.move(xa, box)
        xb, _ = theirs.move(xb, box)
        assert mine.n_accepted() == theirs.n_accepted()
        if precision == np.float32:
            np.testing.assert_array_equal(xa, xb)
        else:
            np.testing.assert_allclose(xa, xb, rtol=0, atol=1e-12)
        np.testing.assert_allclose(mine.last_raw_log_probability(), theirs.last_raw_log_probability(), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(mine.last_log_probability(), theirs.last_log_probability(), rtol=1e-4, atol=1e-4)
    assert mine.n_proposed() == theirs.n_proposed()


def test_tibd_log_probability_matches_oracle():
    """One proposal per move in f64: the reported raw log acceptance equals compute_raw_ratio_given_weights of the Python
    reference (exchange_mover.py:298-323) evaluated with the oracle's weights on the coordinates before / after."""
    o = ops()
    x, params, box, mols, lig = ligand_water_system(120)
    N, radius = len(x), 0.7
    mover = o.TIBDExchangeMove_f64(N, lig, mols, params, TEMP, BETA, CUTOFF, radius, 11, 1, 1)
    vol_inner = 4.0 / 3.0 * np.pi * radius**3
    vol_outer = np.prod(np.diag(box)) - vol_inner
    checked = 0
    for _ in range(300):
        if checked >= 3:
            break
        center = x[lig].mean(0)
        inner, outer = O.water_groups(x, box, center, mols, radius)
        w_before = O.bd_log_weights(x, params, box, mols, BETA, CUTOFF, TEMP)
        acc = mover.n_accepted()
        x_new, _ = mover.move(x, box)
        raw = mover.last_raw_log_probability()
        if mover.n_accepted() > acc:
            new_inner, _ = O.water_groups(x_new, box, center, mols, radius)
            moved = (set(inner) ^ set(new_inner)).pop()
            to_inner = moved in set(new_inner)
            src, dest = (outer, inner) if to_inner else (inner, outer)
            w_after = O.bd_log_weights(x_new, params, box, mols, BETA, CUTOFF, TEMP)
            want = O.tibd_raw_log_probability(
                w_before[src], w_after[np.append(dest, moved)], len(src), len(dest), vol_outer if to_inner else vol_inner,
                vol_inner if to_inner else vol_outer,
            )
            np.testing.assert_allclose(raw, want, rtol=1e-8, atol=1e-7)
            checked += 1
        assert mover.last_log_probability() == min(raw, 0.0)
        x = x_new
    assert checked >= 1


@pytest.mark.parametrize("precision", [np.float64, np.float32])
@pytest.mark.parametrize("radius", [0.2, 2.2])
def test_tibd_empty_region_edge_cases(precision, radius):
    """test_cuda_targeted_insertion_mover.py:419-488 (the buckyball edge cases): one of the two regions starts empty - a
    droplet in a large box with a sphere too small to hold any water, or large enough to hold them all - so the first
    proposals can only go one way; later ones choose by the noise.  Against the compiled reference where present."""
    o = ops()
    x, params, box, mols, lig = ligand_water_system(150)
    box = np.eye(3) * 12.0
    N = len(x)
    inner, outer = klass(o, "inner_and_outer_mols", precision)(lig, x, box, mols, radius)
    assert (len(inner) == 0) if radius < 1.0 else (len(outer) == 0)
    args = (N, lig, mols, params, TEMP, BETA, CUTOFF, radius, 2025, 60, 1)
    mine = klass(o, "TIBDExchangeMove", precision)(*args, batch_size=20)
    single = klass(o, "TIBDExchangeMove", precision)(*args[:-2], 1, 1)
    ref = load_reference_ops()
    theirs = klass(ref, "TIBDExchangeMove", precision)(*args, batch_size=20) if ref is not None else None
    xa, xb, xc = x.copy(), x.copy(), x.copy()
    for _ in range(3):
        xa, _ = mine.move(xa, box)
        for _ in range(60):
            xc, _ = single.move(xc, box)
        np.testing.assert_array_equal(xa, xc)  # batch-size independence through the edge cases
        if theirs is not None:
            xb, _ = theirs.move(xb, box)
            assert mine.n_accepted() == theirs.n_accepted()
            if precision == np.float32:
                np.testing.assert_array_equal(xa, xb)
            else:
                np.testing.assert_allclose(xa, xb, rtol=0, atol=1e-12)
    assert mine.n_accepted() == single.n_accepted()
    assert mine.n_proposed() == single.n_proposed() == 180
    if mine.n_accepted() > 0:
        new_inner, new_outer = klass(o, "inner_and_outer_mols", precision)(lig, xa, box, mols, radius)
        assert len(new_inner) > 0 and len(new_outer) > 0  # the empty region got populated
