"""Small BD + TIBD runs for compute-sanitizer (memcheck / racecheck): all phases, accept and reject paths, f32 and f64.

    compute-sanitizer --tool memcheck python profiles/sanitize_exchange.py      (profiles/r1s3_sanitizer_exchange.txt)
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests.common import water_box  # noqa: E402
from timemachine_b200 import custom_ops as o  # noqa: E402

TEMP, BETA, CUTOFF = 300.0, 2.0, 1.2


def system(n_waters, n_first_waters_as_solute):
    """Water box; the first molecules are not exchange targets (they play the solute / ligand)."""
    s = water_box(n_waters + n_first_waters_as_solute, seed=2023)
    n0 = 3 * n_first_waters_as_solute
    mols = [[n0 + 3 * i, n0 + 3 * i + 1, n0 + 3 * i + 2] for i in range(n_waters)]
    return s["x"], s["params"], s["box"], mols, np.arange(n0, dtype=np.int32)


x, params, box, mols, _ = system(150, 1)
for k in (o.BDExchangeMove_f32, o.BDExchangeMove_f64):
    m = k(len(x), mols, params, TEMP, BETA, CUTOFF, 5, 120, 1, batch_size=32)
    xs = x
    for _ in range(2):
        xs, _ = m.move(xs, box)
    print(k.__name__, m.n_accepted(), m.n_proposed(), m.last_log_probability())
x, params, box, mols, lig = system(300, 4)
for k in (o.TIBDExchangeMove_f32, o.TIBDExchangeMove_f64):
    m = k(len(x), lig, mols, params, TEMP, BETA, CUTOFF, 0.8, 5, 120, 1, batch_size=32)
    xs = x
    for _ in range(2):
        xs, _ = m.move(xs, box)
    print(k.__name__, m.n_accepted(), m.n_proposed(), m.last_log_probability())
