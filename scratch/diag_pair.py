"""Diagnostic: two-atom systems, ours vs the compiled reference, to localise per-pair rounding differences."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests.common import load_reference_ops
from timemachine_b200 import custom_ops as ops

ref = load_reference_ops()
rng = np.random.default_rng(0)
box = np.eye(3) * 4.0
def run(q, eps, w, n=4000, label=""):
    ndiff = 0; worst = 0
    mine = ops.NonbondedAllPairs_f32(2, 2.0, 1.2, None, True, 0.1)
    theirs = ref.NonbondedAllPairs_f32(2, 2.0, 1.2, None, True, 0.1)
    first = None
    for k in range(n):
        d = rng.uniform(0.08, 1.19)
        v = rng.normal(size=3); v /= np.linalg.norm(v)
        x = np.array([[1.0, 1.0, 1.0], [1.0, 1.0, 1.0] + d * v]).astype(np.float32).astype(np.float64)
        p = np.array([[q[0], 0.15, eps[0], w[0]], [q[1], 0.16, eps[1], w[1]]]).astype(np.float32).astype(np.float64)
        a = mine.execute(x, p, box, True, False, True)
        b = theirs.execute(x, p, box, True, False, True)
        if not np.array_equal(a[0], b[0]) or a[2] != b[2]:
            ndiff += 1
            if first is None:
                first = (d, a[0][0] * 2**36, b[0][0] * 2**36, a[2] * 2**36, b[2] * 2**36)
    print(f"{label:30s} differing systems {ndiff}/{n}", "first:", first)

run((3.0, -2.0), (0.0, 0.0), (0.0, 0.0), label="ES only, vanilla")
run((0.0, 0.0), (0.7, 0.9), (0.0, 0.0), label="LJ only (q=0), vanilla")
run((3.0, -2.0), (0.7, 0.9), (0.0, 0.0), label="ES+LJ, vanilla")
run((3.0, -2.0), (0.7, 0.9), (0.2, 0.0), label="ES+LJ, alchemical")
# exclusion kernel
mine = ops.NonbondedExclusions_f32(np.array([[0, 1]], dtype=np.int32), np.array([[1.0, 1.0]]), 2.0, 1.2)
theirs = ref.NonbondedExclusions_f32(np.array([[0, 1]], dtype=np.int32), np.array([[1.0, 1.0]]), 2.0, 1.2)
nd = 0
for k in range(2000):
    d = rng.uniform(0.08, 1.19)
    v = rng.normal(size=3); v /= np.linalg.norm(v)
    x = np.array([[1.0, 1.0, 1.0], [1.0, 1.0, 1.0] + d * v]).astype(np.float32).astype(np.float64)
    p = np.array([[3.0, 0.15, 0.7, 0.0], [-2.0, 0.16, 0.9, 0.0]]).astype(np.float32).astype(np.float64)
    a = mine.execute(x, p, box, True, True, True); b = theirs.execute(x, p, box, True, True, True)
    nd += not (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2])
print("exclusion kernel differing", nd, "/2000")
# does the REFERENCE cancel its own exclusions exactly?
ap = ref.NonbondedAllPairs_f32(2, 2.0, 1.2, None, True, 0.1)
nd = 0
for k in range(2000):
    d = rng.uniform(0.08, 1.19)
    v = rng.normal(size=3); v /= np.linalg.norm(v)
    x = np.array([[1.0, 1.0, 1.0], [1.0, 1.0, 1.0] + d * v]).astype(np.float32).astype(np.float64)
    p = np.array([[3.0, 0.15, 0.7, 0.0], [-2.0, 0.16, 0.9, 0.0]]).astype(np.float32).astype(np.float64)
    a = ap.execute(x, p, box, True, True, True); b = theirs.execute(x, p, box, True, True, True)
    nd += bool(np.any(a[0] + b[0]) or np.any(a[1] + b[1]) or (a[2] + b[2]) != 0)
print("reference all-pairs + reference exclusions not cancelling exactly:", nd, "/2000")
