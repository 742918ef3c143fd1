"""CPU: an executable statement of the index arithmetic of the tile kernel's row-run phase B (csrc/nb_tiles_cq.cuh,
cq_tile_prefilter_runs) - mask layout, exclusive scan, equal contiguous chunks, binary search for a chunk's first row,
select-by-rank inside it, the walk with its flush rule - checked for the properties the kernel relies on: every candidate is
evaluated exactly once, by exactly one lane; lanes differ by at most one step; the flushed register sums add up to the row
sums.  (The kernel itself is tested bitwise against the reference on the GPU; this pins the design.)"""

import numpy as np
import pytest


def column_of(i, b):
    """bit b of row i's mask <-> column position: even offsets first, then odd (phase A files double round R under bits R and
    16 + R: columns i + 2R and i + 2R + 1)"""
    return (i + ((b & 15) << 1) + (b >> 4)) & 31


def popc(v):
    return bin(v & 0xFFFFFFFF).count("1")


def drop_lowest(cur, skip):
    """the kernel's select-by-rank: clear the `skip` lowest set bits with five halvings"""
    pos = 0
    for w in (16, 8, 4, 2, 1):
        cnt = popc((cur >> pos) & ((1 << w) - 1))
        if cnt <= skip:
            skip -= cnt
            pos += w
    return (cur >> pos) << pos


def walk(masks, values):
    """masks[32]: candidate mask per row; values[i][j]: what pair (i, j) contributes.  Returns (visits {(i, j): count},
    row sums from the flushes, steps per lane, number of flushes)."""
    c = [popc(m) for m in masks]
    excl = np.concatenate([[0], np.cumsum(c)[:-1]]).astype(int)
    n = int(sum(c))
    visits, row_sums, steps, flushes = {}, np.zeros(32), [], 0
    if n == 0:
        return visits, row_sums, [0] * 32, 0
    nonempty = sum(1 << i for i in range(32) if c[i])
    K = (n + 31) >> 5
    for lane in range(32):
        e0 = lane * K
        todo = min(K, n - e0)
        steps.append(max(todo, 0))
        if todo <= 0:
            continue
        i = 0
        for step in (16, 8, 4, 2, 1):  # last row whose prefix is <= e0
            if excl[i + step] <= e0:
                i += step
        assert c[i] > 0 and excl[i] <= e0 < excl[i] + c[i]
        cur = drop_lowest(masks[i], e0 - excl[i])
        acc = 0.0
        for k in range(todo):
            assert cur != 0
            b = (cur & -cur).bit_length() - 1
            j = column_of(i, b)
            visits[(i, j)] = visits.get((i, j), 0) + 1
            acc += values[i][j]
            cur &= cur - 1
            if cur == 0 or k == todo - 1:
                row_sums[i] += acc  # the shared-memory atomics of the flush
                acc = 0.0
                flushes += 1
                above = nonempty & ~((2 << i) - 1) & 0xFFFFFFFF
                if above:
                    i = (above & -above).bit_length() - 1
                    cur = masks[i]
    return visits, row_sums, steps, flushes


def random_masks(rng, kind):
    if kind == "empty":
        return [0] * 32
    if kind == "full":
        return [0xFFFFFFFF] * 32
    if kind == "single":
        m = [0] * 32
        m[int(rng.integers(32))] = 1 << int(rng.integers(32))
        return m
    if kind == "one row":
        m = [0] * 32
        m[int(rng.integers(32))] = int(rng.integers(1, 2**32))
        return m
    if kind == "half":  # one half of a split tile: double rounds 8..15 only
        return [int(rng.integers(0, 2**32)) & 0xFF00FF00 for _ in range(32)]
    p = {"sparse": 0.05, "typical": 0.34, "dense": 0.9}[kind]
    return [int(sum(1 << b for b in range(32) if rng.random() < p)) * int(rng.random() < 0.9) for _ in range(32)]


@pytest.mark.parametrize("kind", ["empty", "full", "single", "one row", "half", "sparse", "typical", "dense"])
def test_every_candidate_once_balanced_and_row_sums(kind):
    rng = np.random.default_rng(sum(map(ord, kind)))
    for _ in range(20):
        masks = random_masks(rng, kind)
        values = rng.normal(size=(32, 32))
        visits, row_sums, steps, flushes = walk(masks, values)
        want = {(i, column_of(i, b)) for i in range(32) for b in range(32) if masks[i] >> b & 1}
        assert set(visits) == want and all(v == 1 for v in visits.values())
        n = len(want)
        assert sum(steps) == n and max(steps) == (n + 31) // 32
        busy = [s for s in steps if s]
        K = (n + 31) // 32
        assert all(s == K for s in busy[:-1]) and (not busy or 0 < busy[-1] <= K)  # only the last busy lane may be short
        assert steps == busy + [0] * (32 - len(busy))  # and the idle lanes are the trailing ones
        ref = np.zeros(32)
        for (i, j) in want:
            ref[i] += values[i][j]
        np.testing.assert_allclose(row_sums, ref, atol=1e-12)
        rows_hit = sum(1 for m in masks if m)
        assert flushes <= rows_hit + sum(1 for s in steps if s)  # a flush per row, plus one where a chunk ends inside a row


def test_column_mapping_is_a_bijection_per_row():
    for i in range(32):
        assert sorted(column_of(i, b) for b in range(32)) == list(range(32))
    # phase A: double round R of lane i tests columns i + 2R (bit R) and i + 2R + 1 (bit 16 + R)
    for i in (0, 5, 31):
        for R in range(16):
            assert column_of(i, R) == (i + 2 * R) % 32 and column_of(i, 16 + R) == (i + 2 * R + 1) % 32


def test_drop_lowest_is_select_by_rank():
    rng = np.random.default_rng(1)
    for _ in range(500):
        m = int(rng.integers(1, 2**32))
        k = int(rng.integers(0, popc(m)))
        want = m
        for _ in range(k):
            want &= want - 1
        assert drop_lowest(m, k) == want
