import sys, numpy as np
sys.path.insert(0, '/root/repo')
from tests.common import water_box
from timemachine_b200 import custom_ops as o
s = water_box(10000, seed=11)
x, box, params, N = s["x"], s["box"], s["params"], s["N"]
mols = [[3*i, 3*i+1, 3*i+2] for i in range(10000)]
m = o.BDExchangeMove_f32(N, mols, params, 300.0, 2.0, 1.2, 2024, 1000, 1, batch_size=250)
xs = x
for _ in range(2):
    xs, _ = m.move(xs, box)
