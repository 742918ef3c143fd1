"""GPU: the HREX driver on real Contexts (timemachine_b200/hrex.py).

  * world_size 1 against the reference's algorithm written out with raw custom_ops calls (fe/free_energy.py:1485-1547:
    load replica, set_params, multiple_steps, compute_potential_matrix over all replicas at once);
  * the replica-per-rank layout (2 processes, gloo for the all-gather, both on cuda:0 so the test runs on a one-GPU
    box) against world_size 1: same frames, velocities, energies and permutations bit for bit, WITH thermostat noise -
    the counter-based noise stream makes trajectories independent of the layout."""

import os
import socket
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_STATES, N_FRAMES, STEPS, N_EQ = 3, 3, 20, 10
TEMPERATURE, DT, FRICTION = 300.0, 1.0e-3, 1.0
SEED = 2025


def _system():
    from tests.common import round_to_f32, water_box

    s = water_box(700, seed=3)  # box 2.76 nm >= 2 (cutoff + padding)
    s["x"] = round_to_f32(s["x"])
    return s


def _make(s):
    """Context + parameter sets: the states scale the charges (a stand-in for lambda windows)."""
    from timemachine_b200 import custom_ops, lib, potentials

    N = s["N"]
    nb = potentials.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], 2.0, 1.2)
    pot = potentials.SummedPotential(
        [potentials.HarmonicBond(s["bond_idxs"]), potentials.HarmonicAngle(s["angle_idxs"]), nb],
        [s["bond_params"], s["angle_params"], s["params"]],
    )
    impl = pot.to_gpu(np.float32).unbound_impl
    params_by_state = []
    for k in range(N_STATES):
        p = s["params"].copy()
        p[:, 0] *= 1.0 - 0.01 * k
        params_by_state.append(np.concatenate([s["bond_params"].reshape(-1), s["angle_params"].reshape(-1), p.reshape(-1)]))
    params_by_state = np.array(params_by_state)
    intg = lib.LangevinIntegrator(TEMPERATURE, DT, FRICTION, s["masses"], SEED).impl()
    ctx = custom_ops.Context(s["x"], np.zeros_like(s["x"]), s["box"], intg, [custom_ops.BoundPotential(impl, params_by_state[0])])
    return ctx, params_by_state


def _replicas(s):
    from timemachine_b200 import hrex as H

    rng = np.random.default_rng(1)
    return [H.CoordsVelBox(s["x"] + rng.normal(0, 0.002, s["x"].shape), np.zeros_like(s["x"]), s["box"].copy()) for _ in range(N_STATES)]


def _run(dist=None, out_dir=None, record=None):
    from timemachine_b200 import hrex as H

    s = _system()
    ctx, params_by_state = _make(s)
    sampler = H.ContextSampler(ctx, params_by_state)
    md = H.HREXMDParams(n_frames=N_FRAMES, steps_per_frame=STEPS, n_eq_steps=N_EQ, seed=SEED, max_delta_states=1)
    hook = (lambda f, U, hx: record.append(U.copy())) if record is not None else None
    trajs, diag, hx = H.run_sims_hrex(sampler, _replicas(s), TEMPERATURE, md, out_dir=out_dir, dist=dist, on_iteration=hook)
    return dict(
        history=np.array(diag.replica_idx_by_state_by_iter), frac=np.array(diag.fraction_accepted_by_pair_by_iter),
        final=np.array(hx.replica_idx_by_state), frames=np.array([[f for f in t.frames] for t in trajs]),
        boxes=np.array([t.boxes for t in trajs]), vels=np.array([t.final_velocities for t in trajs]),
    )


def test_sequential_driver_equals_the_reference_algorithm_spelled_out():
    from timemachine_b200 import hrex as H

    Us = []
    got = _run(record=Us)
    assert got["frames"].shape == (N_STATES, N_FRAMES, got["frames"].shape[2], 3) and np.isfinite(got["frames"]).all()

    # the reference's loop, one Context, raw calls
    s = _system()
    ctx, params_by_state = _make(s)
    bp = ctx.get_potentials()[0]
    intg = ctx.get_integrator()
    hx = H.HREX.from_replicas(_replicas(s))
    steps_done = [0] * N_STATES
    frames = [[] for _ in range(N_STATES)]
    history = []
    for frame in range(N_FRAMES):
        reps = list(hx.replicas)
        for state, r in enumerate(hx.replica_idx_by_state):
            ctx.set_x_t(reps[r].coords)
            ctx.set_v_t(reps[r].velocities)
            ctx.set_box(reps[r].box)
            bp.set_params(params_by_state[state])
            n = STEPS + (N_EQ if frame == 0 else 0)
            intg.set_step((r << 40) + steps_done[r])
            xs, boxes = ctx.multiple_steps(n)
            steps_done[r] += n
            reps[r] = H.CoordsVelBox(xs[-1], ctx.get_v_t(), boxes[-1])
            frames[state].append(xs[-1])
        hx = H.HREX(reps, hx.replica_idx_by_state)
        U_kl = H.compute_potential_matrix(bp.get_potential(), hx, params_by_state, 1)  # all replicas in one sparse batch
        np.testing.assert_array_equal(U_kl, Us[frame])
        U_kl = H.verify_and_sanitize_potential_matrix(U_kl, hx.replica_idx_by_state)
        history.append(list(hx.replica_idx_by_state))
        hx, _ = hx.attempt_neighbor_swaps_fast([(0, 1), (1, 2)], -U_kl / (H.BOLTZ * TEMPERATURE), N_STATES**3, SEED + frame + 1)
    np.testing.assert_array_equal(np.array(history), got["history"])
    np.testing.assert_array_equal(np.array(frames), got["frames"])
    # diagonal of U_kl (replica in its own state) is finite, the far corner is not evaluated with max_delta_states = 1
    assert np.isfinite(np.diagonal(Us[0])).all() and np.isinf(Us[0][0, 2])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = _run(dist=dist, out_dir=Path(out_dir) / "traj")
        np.savez(Path(out_dir) / f"rank{rank}.npz", **res)
    finally:
        dist.destroy_process_group()


def test_replica_per_rank_layout_is_bitwise_the_sequential_run(tmp_path):
    import torch.multiprocessing as mp

    ref = _run()
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        got = dict(np.load(tmp_path / f"rank{rank}.npz"))
        for k, v in ref.items():
            np.testing.assert_array_equal(got[k], v, err_msg=f"rank {rank}: {k}")
    # frames of different states differ (the thermostat noise is per replica), and the run moved the atoms
    assert np.abs(ref["frames"][0, 0] - ref["frames"][1, 0]).max() > 1e-4


# ---------------------------------------------------------------------------------------------------------------------
# The layout bench.py runs at --gpus 4 / 8: one replica per rank, resident on the device (DeviceResidentSampler: swaps
# through BoundPotential.set_params_device, U_kl rows from energy-only execute_device calls on the resident coordinates,
# three candidate states on the middle ranks), for many frames with swaps being accepted.
K4_STATES, K4_FRAMES, K4_STEPS = 4, 50, 10


def _run_k4(resident, dist=None, out_dir=None):
    """4 windows whose charges differ by 0.2 % (so that neighbour swaps ARE accepted), 50 frames of 10 steps."""
    import torch

    from timemachine_b200 import custom_ops, lib, potentials
    from timemachine_b200 import hrex as H

    s = _system()
    N = s["N"]
    nb = potentials.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], 2.0, 1.2)
    pot = potentials.SummedPotential(
        [potentials.HarmonicBond(s["bond_idxs"]), potentials.HarmonicAngle(s["angle_idxs"]), nb],
        [s["bond_params"], s["angle_params"], s["params"]],
    )
    impl = pot.to_gpu(np.float32).unbound_impl
    params_by_state = []
    for k in range(K4_STATES):
        p = s["params"].copy()
        p[:, 0] *= 1.0 - 0.002 * k
        params_by_state.append(np.concatenate([s["bond_params"].reshape(-1), s["angle_params"].reshape(-1), p.reshape(-1)]))
    params_by_state = np.array(params_by_state)
    intg = lib.LangevinIntegrator(TEMPERATURE, DT, FRICTION, s["masses"], SEED).impl()
    bp = custom_ops.BoundPotential(impl, params_by_state[0])
    ctx = custom_ops.Context(s["x"], np.zeros_like(s["x"]), s["box"], intg, [bp])
    sampler = H.DeviceResidentSampler(ctx, params_by_state, torch.device("cuda", 0)) if resident else H.ContextSampler(ctx, params_by_state)
    rng = np.random.default_rng(1)
    replicas = [H.CoordsVelBox(s["x"] + rng.normal(0, 0.002, s["x"].shape), np.zeros_like(s["x"]), s["box"].copy()) for _ in range(K4_STATES)]
    md = H.HREXMDParams(n_frames=K4_FRAMES, steps_per_frame=K4_STEPS, n_eq_steps=0, seed=SEED, max_delta_states=1)
    Us = []
    trajs, diag, hx = H.run_sims_hrex(sampler, replicas, TEMPERATURE, md, out_dir=out_dir, dist=dist, on_iteration=lambda f, U, h: Us.append(U.copy()))
    # what the BoundPotential holds now must be, bit for bit, the parameter set of the state its last replica was sampled in
    return dict(
        history=np.array(diag.replica_idx_by_state_by_iter), final=np.array(hx.replica_idx_by_state),
        frames=np.array([[f for f in t.frames] for t in trajs]), vels=np.array([t.final_velocities for t in trajs]), U=np.array(Us),
    )


def _worker_k4(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = _run_k4(True, dist=dist, out_dir=Path(out_dir) / "traj")
        np.savez(Path(out_dir) / f"rank{rank}.npz", **res)
    finally:
        dist.destroy_process_group()


def test_four_resident_replicas_with_accepted_swaps_equal_the_sequential_host_run(tmp_path):
    import torch.multiprocessing as mp

    ref = _run_k4(False)  # world_size 1, the reference's host-buffer algorithm (set_x_t / set_params every frame)
    assert np.isfinite(ref["frames"]).all() and np.isfinite(ref["vels"]).all()
    n_perms = len({tuple(p) for p in ref["history"]})
    assert n_perms >= 4, f"only {n_perms} distinct permutations in 50 frames: swaps are not being accepted, the test proves nothing"
    # every replica has been sampled under at least two states (the swap path of the resident sampler is exercised)
    assert all(len(set(ref["history"][:, s])) >= 2 for s in range(K4_STATES))
    # one process, every replica loaded in turn into the resident sampler (the capture-to-host fall-back path)
    seq = _run_k4(True, out_dir=tmp_path / "seq")
    for k, v in ref.items():
        np.testing.assert_array_equal(seq[k], v, err_msg=f"resident sampler, one process: {k}")
    mp.spawn(_worker_k4, args=(4, _free_port(), str(tmp_path)), nprocs=4, join=True)
    for rank in range(4):
        got = dict(np.load(tmp_path / f"rank{rank}.npz"))
        for k, v in ref.items():
            np.testing.assert_array_equal(got[k], v, err_msg=f"rank {rank} of 4: {k}")


def test_set_params_device_is_bitwise_set_params():
    """BoundPotential.set_params_device (reference bound_potential.cu:139-147) against set_params: same forces and energy
    bit for bit, on the stream the copy was enqueued on; a larger size than the buffer is refused with the reference's text."""
    import torch

    from timemachine_b200 import custom_ops, potentials

    s = _system()
    N = s["N"]
    nb = potentials.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], 2.0, 1.2).to_gpu(np.float32).unbound_impl
    p0, p1 = s["params"].copy(), s["params"].copy()
    p1[:, 0] *= 0.9
    p1[::7, 3] = 0.3
    a = custom_ops.BoundPotential(nb, p0)
    b = custom_ops.BoundPotential(nb, p0)
    a.set_params(p1)
    d_p1 = torch.from_numpy(p1.reshape(-1)).cuda()
    stream = torch.cuda.Stream()
    b.set_params_device(d_p1.data_ptr(), p1.size, stream.cuda_stream)
    stream.synchronize()
    ra = a.execute(s["x"], s["box"])
    rb = b.execute(s["x"], s["box"])
    np.testing.assert_array_equal(ra[0], rb[0])
    assert ra[1] == rb[1]
    with pytest.raises(RuntimeError, match="parameter size is greater than device buffer size"):
        b.set_params_device(d_p1.data_ptr(), p1.size + 4, stream.cuda_stream)
