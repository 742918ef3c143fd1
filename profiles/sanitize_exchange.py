"""Small BD + TIBD runs for compute-sanitizer (memcheck / racecheck): all phases, accept and reject paths, f32 and f64."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests.test_exchange_gpu import BETA, CUTOFF, TEMP, ligand_water_system, water_system  # noqa: E402
from timemachine_b200 import custom_ops as o  # noqa: E402

x, params, box, mols = water_system(150, ions=3)
for k in (o.BDExchangeMove_f32, o.BDExchangeMove_f64):
    m = k(len(x), mols, params, TEMP, BETA, CUTOFF, 5, 120, 1, batch_size=32)
    xs = x
    for _ in range(2):
        xs, _ = m.move(xs, box)
    print(k.__name__, m.n_accepted(), m.n_proposed(), m.last_log_probability())
x, params, box, mols, lig = ligand_water_system(300)
for k in (o.TIBDExchangeMove_f32, o.TIBDExchangeMove_f64):
    m = k(len(x), lig, mols, params, TEMP, BETA, CUTOFF, 0.8, 5, 120, 1, batch_size=32)
    xs = x
    for _ in range(2):
        xs, _ = m.move(xs, box)
    print(k.__name__, m.n_accepted(), m.n_proposed(), m.last_log_probability())
