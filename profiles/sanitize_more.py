"""compute-sanitizer driver, second part: the f64 (rotation) tile kernel, the list build growing its buffer, NPT moves, local MD.

    TMB_NBLIST_TILES_PER_ATOM_X100=5 compute-sanitizer --tool memcheck  python profiles/sanitize_more.py
    compute-sanitizer --tool racecheck python profiles/sanitize_more.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from timemachine_b200 import custom_ops as ops  # noqa: E402
from timemachine_b200 import potentials as P  # noqa: E402

s = bench.build_system(900, 24, seed=7)
N = s["N"]
flat = bench.flat_params(s, 0.5)
# f64 potentials: rotation tile kernel, all outputs
impl64 = bench.make_potential(P, s).to_gpu(np.float64).unbound_impl
du_dx, du_dp, u = impl64.execute(s["x"], flat, s["box"])
assert np.isfinite(u)
# f32 + barostat + MD (a tiny initial tile buffer, when selected by the environment, makes the first call grow it)
impl = bench.make_potential(P, s).to_gpu(np.float32).unbound_impl
bp = ops.BoundPotential(impl, flat)
groups = [np.arange(i, i + 3, dtype=np.int32) for i in range(0, s["n_env"], 3)] + [s["lig_idx"].astype(np.int32)]
baro = ops.MonteCarloBarostat(N, 1.013, 300.0, groups, 5, [bp], 99, True, 0.0)
intg = ops.LangevinIntegrator(s["masses"], 300.0, 5e-4, 20.0, 1)
ctx = ops.Context(s["x"], np.zeros_like(s["x"]), s["box"], intg, [bp], movers=[baro])
for attempt in range(2):
    try:
        xs, boxes = ctx.multiple_steps(30)
        break
    except RuntimeError as e:  # tile buffer grown: restore and go again
        assert "overflow" in str(e)
        ctx.set_x_t(s["x"]); ctx.set_v_t(np.zeros_like(s["x"])); ctx.set_box(s["box"])
assert np.isfinite(xs).all()
# local MD around a ligand atom
xs, _ = ctx.multiple_steps_local(20, s["lig_idx"][:4].astype(np.int32), radius=0.8, k=1000.0, seed=3)
assert np.isfinite(xs).all()
print("ok", N, "atoms")
