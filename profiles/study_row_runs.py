"""Design study (CPU): column-side ATOMS passes per phase-B step of the row-run tile path for different orders in which a
row's candidates are walked.  Real tile structure (Hilbert-sorted 30 000-atom water box, the kernel's tiles), the kernel's
chunk assignment: lane l takes entries [l K, (l + 1) K) of the row-sorted candidate list, K = ceil(n / 32).

    python profiles/study_row_runs.py
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from scipy.spatial import cKDTree

from oracle import tm_oracle as O
from tests.common import water_box

s = water_box(10000, seed=2022, jitter=0.01)
x, box = s["x"], s["box"]
L = box[0, 0]
N = len(x)
cutoff, pad = 1.2, 0.1
perm = O.hilbert_perm(x, box)
xs = x[perm] % L
tree = cKDTree(xs, boxsize=L)
rng = np.random.default_rng(0)
nb = (N + 31) // 32

ORDERS = {
    # offset sequence in which row i walks the column positions (i + offset) % 32
    "even offsets then odd (production)": [2 * (b & 15) + (b >> 4) for b in range(32)],
    "natural 0..31": list(range(32)),
    "bit reversed": [int(f"{b:05b}"[::-1], 2) for b in range(32)],
    "stride 5": [(5 * b) % 32 for b in range(32)],
    "stride 11": [(11 * b) % 32 for b in range(32)],
    "absolute column 0..31 (no rotation)": None,
}


def walk(Ht, order, alternate=False):
    """per step: (column passes, row-flush passes, active lanes)"""
    rows = []
    for i in range(32):
        if order is None:
            js = [j for j in range(32) if Ht[i, j]]
        else:
            js = [(i + o) % 32 for o in order if Ht[i, (i + o) % 32]]
        rows.append(js)
    ent = [(i, j) for i in range(32) for j in rows[i]]
    n = len(ent)
    if n == 0:
        return []
    K = (n + 31) // 32
    out = []
    for k in range(K):
        cols, act = [], 0
        for l in range(32):
            e = l * K + k
            if e < min(n, (l + 1) * K):
                cols.append(ent[e][1])
                act += 1
        out.append((np.bincount(cols, minlength=32).max(), act))
    return out


res = {k: [] for k in ORDERS}
for rb in rng.choice(nb - 1, 60, replace=False):
    ra = np.arange(rb * 32, rb * 32 + 32)
    nbrs = tree.query_ball_point(xs[ra], cutoff + pad)
    cols = np.unique(np.concatenate([np.array(v) for v in nbrs]))
    cols = cols[cols >= rb * 32 + 32]
    hit = tree.query_ball_point(xs[ra], cutoff + 0.008)  # the prefilter's candidates
    cidx = {c: k for k, c in enumerate(cols)}
    ntile = (len(cols) + 31) // 32
    H = np.zeros((32, ntile * 32), bool)
    for i, h in enumerate(hit):
        for c in h:
            if c in cidx:
                H[i, cidx[c]] = True
    for t in range(ntile):
        Ht = H[:, t * 32 : (t + 1) * 32]
        for name, order in ORDERS.items():
            res[name] += walk(Ht, order)
for name, v in res.items():
    v = np.array(v)
    print(f"{name:40s} column passes per step {v[:, 0].mean():.2f}   lanes active {v[:, 1].mean():.1f}")

# ---- second part: the family the kernel can stage without bank conflicts: row i walks column (s i + pi(b)) % 32, s odd
print()


def br4(v):
    return int(f"{v:04b}"[::-1], 2)


PIS = {
    "evenodd": [2 * (b & 15) + (b >> 4) for b in range(32)],
    "halves": list(range(32)),
    "halves, rounds bit-reversed": [br4(b & 15) + 16 * (b >> 4) for b in range(32)],
    "evenodd, rounds bit-reversed": [2 * br4(b & 15) + (b >> 4) for b in range(32)],
}
tiles = []
rng = np.random.default_rng(0)
for rb in rng.choice(nb - 1, 60, replace=False):
    ra = np.arange(rb * 32, rb * 32 + 32)
    nbrs = tree.query_ball_point(xs[ra], cutoff + pad)
    cols = np.unique(np.concatenate([np.array(v) for v in nbrs]))
    cols = cols[cols >= rb * 32 + 32]
    hit = tree.query_ball_point(xs[ra], cutoff + 0.008)
    cidx = {c: k for k, c in enumerate(cols)}
    ntile = (len(cols) + 31) // 32
    H = np.zeros((32, ntile * 32), bool)
    for i, h in enumerate(hit):
        for c in h:
            if c in cidx:
                H[i, cidx[c]] = True
    for t in range(ntile):
        tiles.append(H[:, t * 32 : (t + 1) * 32].copy())


def passes_for(s, pi):
    tot, cnt = 0, 0
    for Ht in tiles:
        ent = [(i, (s * i + o) % 32) for i in range(32) for o in pi if Ht[i, (s * i + o) % 32]]
        n = len(ent)
        if n == 0:
            continue
        K = (n + 31) // 32
        for k in range(K):
            cols_k = [ent[l * K + k][1] for l in range(32) if l * K + k < min(n, (l + 1) * K)]
            tot += np.bincount(cols_k, minlength=32).max()
            cnt += 1
    return tot / cnt


for pname, pi in PIS.items():
    print(f"{pname:32s}", "  ".join(f"s={s}: {passes_for(s, pi):.2f}" for s in (1, 3, 5, 7, 9, 11, 13, 15, 17, 21, 25, 29, 31)))
