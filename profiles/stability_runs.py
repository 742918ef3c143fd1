"""Long single-window runs of the bench system at the lambdas that went unstable in 4/8-window runs."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from timemachine_b200 import custom_ops as ops, potentials as P

s = B.build_system(10000, 60, seed=2022)
N = s["N"]
impl = B.make_potential(P, s).to_gpu(np.float32).unbound_impl
kT = 0.008314462618 * B.TEMPERATURE
for lam in (1.0, 6 / 7, 2 / 3, 0.0):
    flat = B.flat_params(s, lam)
    x, v = B.equilibrate(ops, impl, flat, s, seed=107)
    ctx = ops.Context(x, v, s["box"], ops.LangevinIntegrator(s["masses"], B.TEMPERATURE, B.DT, B.FRICTION, 1234), [ops.BoundPotential(impl, flat)])
    t0 = time.time()
    worst = 0.0
    for block in range(25):
        ctx.multiple_steps(4000)
        vv = ctx.get_v_t()
        T = float(np.sum(s["masses"][:, None] * vv * vv)) / (3 * N * kT) * B.TEMPERATURE
        lig = ctx.get_x_t()[s["n_env"]:]
        a, b, c = lig[:-2], lig[1:-1], lig[2:]
        ang = np.degrees(np.arccos(np.clip(np.sum((a - b) * (c - b), 1) / (np.linalg.norm(a - b, axis=1) * np.linalg.norm(c - b, axis=1)), -1, 1)))
        worst = max(worst, ang.max())
    print(f"lambda {lam:.3f}: 100000 steps ok in {time.time()-t0:.1f} s, T = {T:.1f} K, largest ligand angle seen {worst:.1f} deg", flush=True)
