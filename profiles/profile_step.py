"""Small driver for ncu: a few eager MD steps of the bench system (graphs off so every kernel is a plain launch).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python profiles/profile_step.py
    ncu --set full --clock-control none --import-source on -k regex:k_nb_tiles -s 4 -c 2 -o gpurun_out/nb_tiles python profiles/profile_step.py
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from timemachine_b200 import custom_ops as ops  # noqa: E402
from timemachine_b200 import potentials as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--waters", type=int, default=10000)
args = ap.parse_args()

s = bench.build_system(args.waters, 60, seed=2022)
impl = bench.make_potential(P, s).to_gpu(np.float32).unbound_impl
flat = bench.flat_params(s, 0.5)
intg = ops.LangevinIntegrator(s["masses"], 300.0, 2e-4, 50.0, 1)
ctx = ops.Context(s["x"], np.zeros_like(s["x"]), s["box"], intg, [ops.BoundPotential(impl, flat)])
ctx.set_use_graphs(False)
ctx.multiple_steps(args.steps, args.steps + 1)
print("tiles", impl.get_potentials()[3].get_potentials()[0].get_tile_count(), "atoms", s["N"])
