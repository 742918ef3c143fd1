"""Replays rank 2 of the 4-window bench (the run that went unstable under the driver) in ONE process and looks at
the step where it blows up: forces of this repo against the compiled reference on the same coordinates."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
import torch
from timemachine_b200 import custom_ops as ops, potentials as P, replica
from tests.common import load_reference_ops

rank, world = 2, 4
# state held by rank 2 DURING frame f (from the logged run): starts at 2, then the logged post-swap states
seq = [2, 3, 2, 3, 3, 2, 3, 2, 3, 3, 2, 2, 3, 2, 3, 2, 3, 3, 2, 3, 2, 3, 2, 3]
s = B.build_system(10000, 60, seed=2022)
N = s["N"]
lambdas = np.linspace(0, 1, world)
flats = [B.flat_params(s, float(l)) for l in lambdas]
pot = B.make_potential(P, s)
impl = pot.to_gpu(np.float32).unbound_impl
x_eq, v_eq = B.equilibrate(ops, impl, flats[rank], s, seed=100 + rank)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
bp = ops.BoundPotential(impl, flats[rank])
intg = ops.LangevinIntegrator(s["masses"], B.TEMPERATURE, B.DT, B.FRICTION, 1234 + rank)
ctx = ops.Context(x_eq, v_eq, s["box"], intg, [bp])
ctx.set_stream(stream.cuda_stream)
d_x, d_v, d_box = ctx.device_state()
d_params = [torch.from_numpy(f).to(dev) for f in flats]
d_u = torch.zeros(8, dtype=torch.int64, device=dev)
P_total = flats[0].size
kT = 0.008314462618 * B.TEMPERATURE

def temp():
    vv = ctx.get_v_t()
    return float(np.sum(s["masses"][:, None] * vv * vv)) / (3 * N * kT) * B.TEMPERATURE

def exchange(mine, nxt):
    cand = replica.candidate_states(mine, world)
    for slot, k in enumerate(cand):
        impl.execute_device(N, P_total, d_x, d_params[k].data_ptr(), d_box, 0, 0, d_u.data_ptr() + 16 * slot, stream.cuda_stream)
    with torch.cuda.stream(stream):
        host = d_u[: 2 * len(cand)].cpu().numpy()
    if nxt != mine:
        bp.set_params_device(d_params[nxt].data_ptr(), P_total, stream.cuda_stream)
    return [replica.i128_to_energy(int(host[2 * i]), int(host[2 * i + 1])) for i in range(len(cand))]

for f in range(len(seq) - 1):
    ctx.multiple_steps(400, 401)
    row = exchange(seq[f], seq[f + 1])
    xx = ctx.get_x_t()
    print(f"frame {f} state {seq[f]} -> {seq[f+1]} row {row} T {temp():.1f} xrange {xx.min():.2f} {xx.max():.2f}", flush=True)
x0, v0 = ctx.get_x_t(), ctx.get_v_t()
np.savez("gpurun_out/blowup_state.npz", x=x0, v=v0, box=s["box"], state=seq[-1])
ref = load_reference_ops()
rimpl = B.make_reference_potential(ref, s) if ref is not None else None
cur = flats[seq[-1]]

def forces(xx):
    du, _, u = impl.execute(xx, cur, s["box"], True, False, True)
    return du, u

def rforces(xx):
    du, _, u = rimpl.execute(xx, cur, s["box"], True, False, True)
    return du, u

# the blowing frame, 10 steps at a time
prevx, prevv = x0, v0
done = 0
bad_at = None
while done < 400:
    ctx.multiple_steps(10, 11)
    done += 10
    xx, vv = ctx.get_x_t(), ctx.get_v_t()
    vmax = np.abs(vv).max()
    T = temp()
    print(f"step {done}: T {T:.1f} vmax {vmax:.2f} atom {np.unravel_index(np.abs(vv).argmax(), vv.shape)}", flush=True)
    if vmax > 30 or not np.isfinite(vmax):
        bad_at = done
        break
    prevx, prevv = xx, vv
print("bad_at", bad_at)
if bad_at is not None:
    # restart from 10 steps before and go one step at a time (same noise stream position is not restored: friction noise is
    # tiny on this time scale, the point is which force explodes)
    ctx2 = ops.Context(prevx, prevv, s["box"], ops.LangevinIntegrator(s["masses"], B.TEMPERATURE, B.DT, B.FRICTION, 99), [ops.BoundPotential(impl, cur)])
    for k in range(15):
        xx = ctx2.get_x_t()
        du, u = forces(xx)
        fn = np.linalg.norm(du, axis=1)
        a = int(fn.argmax())
        msg = f"  sub {k}: u {u:.2f} fmax {fn[a]:.1f} atom {a} (env {a < s['n_env']}, in mol pos {a % 3 if a < s['n_env'] else a - s['n_env']}) vmax {np.abs(ctx2.get_v_t()).max():.2f}"
        if rimpl is not None:
            rdu, ru = rforces(xx)
            msg += f" | ref u {ru:.2f} fmax {np.linalg.norm(rdu, axis=1).max():.1f} maxdiff {np.abs(rdu - du).max():.3e} ndiff {int((rdu != du).any(axis=1).sum())}"
        # nearest neighbours of the hot atom
        d = xx - xx[a]
        L = s["box"][0, 0]
        d -= L * np.round(d / L)
        r = np.linalg.norm(d, axis=1)
        nn = np.argsort(r)[1:5]
        msg += f" nn {[(int(j), round(float(r[j]), 4)) for j in nn]}"
        print(msg, flush=True)
        ctx2.multiple_steps(1, 2)
