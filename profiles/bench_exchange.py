"""Time one water-exchange move (BDExchangeMove_f32.move on host arrays, the reference's production settings:
1000 proposals per move in batches of 250, timemachine/fe/free_energy.py:130-131) for this repo and, when present, the
compiled reference (oracle/_ref).  Wall clock around the blocking call, median of `reps` after warm-up.

Measurement script (like bench.py's `reference_gpu` leg): the compiled reference is only the thing timed next to this repo,
nothing under timemachine_b200/ depends on it.

    python profiles/bench_exchange.py [n_waters] [proposals] [batch] [reps]
"""

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from tests.common import load_reference_ops, water_box  # noqa: E402


def main():
    n_waters = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    proposals = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 250
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 7
    from timemachine_b200 import custom_ops as mine

    s = water_box(n_waters, seed=11)
    x, box, params, N = s["x"], s["box"], s["params"], s["N"]
    mols = [[3 * i, 3 * i + 1, 3 * i + 2] for i in range(n_waters)]
    out = {"n_atoms": N, "proposals_per_move": proposals, "batch_size": batch}
    impls = {"this_repo": mine}
    ref = load_reference_ops()
    if ref is not None:
        impls["reference"] = ref
    coords = {}
    for name, o in impls.items():
        mover = o.BDExchangeMove_f32(N, mols, params, 300.0, 2.0, 1.2, 2024, proposals, 1, batch_size=batch)
        xs = x
        times = []
        for r in range(reps + 2):
            t0 = time.perf_counter()
            xs, _ = mover.move(xs, box)
            times.append(time.perf_counter() - t0)
        coords[name] = xs
        out[name] = {
            "ms_per_move": 1e3 * float(np.median(times[2:])), "us_per_proposal": 1e6 * float(np.median(times[2:])) / proposals,
            "accepted": mover.n_accepted(), "proposed": mover.n_proposed(),
        }
    if "reference" in coords:
        out["same_coordinates_as_reference"] = bool(np.array_equal(coords["this_repo"], coords["reference"]))
        out["speedup"] = out["reference"]["ms_per_move"] / out["this_repo"]["ms_per_move"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
