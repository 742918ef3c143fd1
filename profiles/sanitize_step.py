"""compute-sanitizer driver: a few MD steps + single evaluations of a small solvated system through every tile-kernel
variant (production, TMB_NB_ASYNC=1, TMB_FUSE_PREPARE=1 select the others from the environment).

    compute-sanitizer --tool memcheck  python profiles/sanitize_step.py
    compute-sanitizer --tool racecheck python profiles/sanitize_step.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from timemachine_b200 import custom_ops as ops  # noqa: E402
from timemachine_b200 import potentials as P  # noqa: E402

s = bench.build_system(900, 24, seed=7)  # ~2.7k atoms, box 3.0 nm
impl = bench.make_potential(P, s).to_gpu(np.float32).unbound_impl
flat = bench.flat_params(s, 0.5)
du_dx, du_dp, u = impl.execute(s["x"], flat, s["box"])
assert np.isfinite(u) and np.isfinite(du_dx).all()
intg = ops.LangevinIntegrator(s["masses"], 300.0, 5e-4, 20.0, 1)
ctx = ops.Context(s["x"], np.zeros_like(s["x"]), s["box"], intg, [ops.BoundPotential(impl, flat)])
xs, _ = ctx.multiple_steps(40)  # graph blocks (fused prepare when selected) + eager steps
assert np.isfinite(xs).all()
print("ok", s["N"], "atoms, u =", u)
