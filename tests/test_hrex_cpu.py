"""CPU: the HREX host driver (timemachine_b200/hrex.py).

  * its pure host functions against the reference's own, executed from the reference source by
    tests/golden/make_golden_hrex.py -> hrex_driver.npz (diagnostics, U_kl index bookkeeping, sanitising);
  * the replica-per-rank layout against the sequential one: world_size 1, 2 and 4 over gloo must produce the same
    permutation history, acceptance counts and frames, bit for bit (the toy sampler is deterministic in
    (replica, steps done), like the Philox-keyed Langevin integrator on the GPU);
  * the StoredArrays chunk layout.
"""

import os
import socket
import types
import warnings
from pathlib import Path

import numpy as np
import pytest

from timemachine_b200 import hrex as H

GOLD = np.load(Path(__file__).parent / "golden" / "hrex_driver.npz")


def test_diagnostics_match_reference_golden():
    hist = GOLD["history"]
    np.testing.assert_array_equal(H.get_cumulative_replica_state_counts(hist), GOLD["counts"])
    np.testing.assert_array_equal(H.estimate_transition_matrix(hist), GOLD["transition_matrix"])
    np.testing.assert_allclose(H.estimate_relaxation_time(GOLD["transition_matrix"]), GOLD["relaxation_time"], rtol=1e-12)
    np.testing.assert_allclose(H.get_normalized_kl_divergence(hist), GOLD["kl"], rtol=1e-12)
    n_iters, n_states = hist.shape
    samples = [[it * 100 + s for s in range(n_states)] for it in range(n_iters)]
    np.testing.assert_array_equal(np.array(H.get_samples_by_iter_by_replica(samples, hist.tolist())), GOLD["by_replica_codes"])
    d = H.HREXDiagnostics(hist.tolist(), GOLD["frac"].tolist())
    np.testing.assert_array_equal(d.cumulative_swap_acceptance_rates, GOLD["cum_rates"])
    np.testing.assert_allclose(d.relaxation_time, GOLD["relaxation_time"], rtol=1e-12)
    np.testing.assert_array_equal([H.get_swap_attempts_per_iter_heuristic(k) for k in range(1, 10)], GOLD["swap_heuristic"])


def test_potential_matrix_bookkeeping_matches_reference_golden():
    coords, boxes, params = GOLD["coords"], GOLD["boxes"], GOLD["params"]
    n = len(coords)

    def energy(ci, pi):
        return coords[ci].sum() * 10 + params[pi].sum() + boxes[ci][0, 0]

    class FakePotential:
        def execute_batch_sparse(self, cs, ps, bs, cidx, pidx, dx, dp, du):
            assert (dx, dp, du) == (False, False, True)
            assert cidx.dtype == np.uint32 and pidx.dtype == np.uint32
            return None, None, np.array([energy(c, p) for c, p in zip(cidx, pidx)])

        def execute_batch(self, cs, ps, bs, dx, dp, du):
            return None, None, np.array([[energy(c, p) for p in range(len(ps))] for c in range(len(cs))])

    hx = H.HREX([H.CoordsVelBox(coords[i], None, boxes[i]) for i in range(n)], GOLD["replica_idx_by_state"].tolist())
    for k in (1, 2, None):
        np.testing.assert_array_equal(H.compute_potential_matrix(FakePotential(), hx, params, k), GOLD[f"U_kl_k{k}"])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        clean = H.verify_and_sanitize_potential_matrix(GOLD["dirty"], GOLD["replica_idx_by_state"].tolist())
        assert any(issubclass(x.category, H.IndeterminateEnergyWarning) for x in w)
    np.testing.assert_array_equal(clean, GOLD["clean"])
    bad = GOLD["U_kl_k2"].copy()
    r0 = GOLD["replica_idx_by_state"][0]
    bad[r0, 0] = np.nan
    with pytest.raises(AssertionError, match="non-finite"):
        H.verify_and_sanitize_potential_matrix(bad, GOLD["replica_idx_by_state"].tolist())


def test_stored_arrays_layout(tmp_path):
    sa = H.StoredArrays(tmp_path / "s")
    sa.extend([np.array([1, 2, 3]), np.array([4, 5, 6])])
    sa.extend([np.array([7, 8, 9])])
    assert len(sa) == 3 and [x.tolist() for x in sa] == [[1, 2, 3], [4, 5, 6], [7, 8, 9]]
    assert sa[2].tolist() == [7, 8, 9] and sa[-3].tolist() == [1, 2, 3]
    assert sorted(p.name for p in (tmp_path / "s").iterdir()) == ["0.npy", "1.npy"]  # reference: <prefix>/<idx>.npy
    again = H.StoredArrays.load(tmp_path / "s")
    assert again == sa
    with pytest.raises(NotImplementedError):
        sa[0:2]
    tmp = H.StoredArrays()
    tmp.extend([np.zeros(2)])
    assert len(tmp) == 1


def test_hrex_bookkeeping():
    hx = H.HREX.from_replicas(["a", "b", "c"])
    assert hx.replica_idx_by_state == [0, 1, 2]
    hx = H.HREX(["a", "b", "c"], [2, 0, 1])
    assert hx.state_replica_pairs == [(0, "c"), (1, "a"), (2, "b")]
    hx2, samples = hx.sample_replicas(lambda rep, s: f"{rep}{s}", lambda smp: smp.upper())
    assert samples == ["c0", "a1", "b2"] and hx2.replicas == ["A1", "B2", "C0"] and hx2.replica_idx_by_state == [2, 0, 1]
    # strongly favourable swap is taken, the permutation stays a permutation, counts add up
    log_q = -np.array([[10.0, 0.0, 50.0], [0.0, 10.0, 50.0], [50.0, 50.0, 0.0]])
    hx3, frac = H.HREX.from_replicas([0, 1, 2]).attempt_neighbor_swaps_fast([(0, 1), (1, 2)], log_q, 7, seed=3)
    assert sorted(hx3.replica_idx_by_state) == [0, 1, 2] and sum(p for _, p in frac) == 7
    assert hx3.replica_idx_by_state[2] == 2  # replica 2 never leaves state 2 (swap costs 100 kT)
    # deterministic in the seed
    assert H.HREX.from_replicas([0, 1, 2]).attempt_neighbor_swaps_fast([(0, 1), (1, 2)], log_q, 7, seed=3)[0] == hx3


# ---------------------------------------------------------------------------------------------------------------------
class ToySampler:
    """Particles in state-dependent harmonic wells; 'MD' is a deterministic contraction plus counter-keyed noise, so a
    replica's trajectory depends only on (replica, steps done, state history) - never on which rank ran it."""

    def __init__(self, n_states):
        self.centers = np.linspace(0.0, 1.0, n_states)
        self.k = 3.0

    def sample(self, xvb, replica_idx, state_idx, steps_done, n_steps):
        rng = np.random.Generator(np.random.Philox(key=[replica_idx, steps_done]))
        x = xvb.coords
        for _ in range(n_steps):
            x = x + 0.2 * (self.centers[state_idx] - x) + 0.08 * rng.normal(size=x.shape)
        # a pretend water sampler: counts keyed like the trajectory, so they cannot depend on the layout either
        self.water_sampling_counts = {state_idx: (int(rng.integers(0, 5)), 10 * n_steps)}
        return H.CoordsVelBox(x, xvb.velocities + 1.0, xvb.box), None

    def energies(self, xvbs, cidx, pidx):
        return np.array([0.5 * self.k * np.sum((xvbs[c].coords - self.centers[p]) ** 2) for c, p in zip(cidx, pidx)])


def _toy_run(n_states, dist=None, out_dir=None):
    rng = np.random.default_rng(7)
    replicas = [H.CoordsVelBox(rng.normal(size=(5, 3)) * 0.1 + c, np.zeros((5, 3)), np.eye(3) * (2 + i)) for i, c in enumerate(np.linspace(0, 1, n_states))]
    params = H.HREXMDParams(n_frames=12, steps_per_frame=3, n_eq_steps=2, seed=11, max_delta_states=2)
    return H.run_sims_hrex(ToySampler(n_states), replicas, 300.0, params, out_dir=out_dir, dist=dist)


def _summarise(trajs, diag, hx):
    return dict(
        history=np.array(diag.replica_idx_by_state_by_iter), frac=np.array(diag.fraction_accepted_by_pair_by_iter),
        final=np.array(hx.replica_idx_by_state), frames=np.array([[f for f in t.frames] for t in trajs]),
        boxes=np.array([t.boxes for t in trajs]), vels=np.array([t.final_velocities for t in trajs]),
        water=diag.water_sampling_diagnostics.proposals_by_state_by_iter,
    )


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_states, out_dir):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = _summarise(*_toy_run(n_states, dist=dist, out_dir=Path(out_dir) / "traj"))
        np.savez(Path(out_dir) / f"rank{rank}.npz", **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_states", [(2, 4), (4, 6)])
def test_replica_per_rank_layout_equals_sequential(tmp_path, world, n_states):
    import torch.multiprocessing as mp

    ref = _summarise(*_toy_run(n_states))  # world_size 1: the reference's sequential algorithm
    assert ref["history"].shape == (12, n_states) and ref["frames"].shape == (n_states, 12, 5, 3)
    assert ref["water"].shape == (12, n_states, 2) and np.all(ref["water"][1:, :, 1] == 30) and np.all(ref["water"][0, :, 1] == 50)
    assert len({tuple(p) for p in ref["history"]}) > 1, "no swap was ever accepted: the test would prove nothing"
    mp.spawn(_worker, args=(world, _free_port(), n_states, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        got = dict(np.load(tmp_path / f"rank{rank}.npz"))
        for k, v in ref.items():
            np.testing.assert_array_equal(got[k], v, err_msg=f"rank {rank}: {k}")


def test_two_state_identity_move_and_context_sampler_contract():
    trajs, diag, hx = _toy_run(2)
    assert np.array(diag.fraction_accepted_by_pair_by_iter).shape == (12, 1, 2)  # identity pair removed from the stats
    # ContextSampler only needs these members of a Context
    calls = []
    bp = types.SimpleNamespace(get_potential=lambda: "pot", set_params=lambda p: calls.append(("params", p.tolist())))
    intg = types.SimpleNamespace(set_step=lambda s: calls.append(("step", s)))
    ctx = types.SimpleNamespace(
        get_potentials=lambda: [bp], get_integrator=lambda: intg, get_barostat=lambda: None,
        set_x_t=lambda x: calls.append("x"), set_v_t=lambda v: calls.append("v"), set_box=lambda b: calls.append("box"),
        multiple_steps=lambda n: (np.zeros((1, 2, 3)), np.eye(3)[None]), get_v_t=lambda: np.ones((2, 3)),
    )
    s = H.ContextSampler(ctx, np.arange(6.0).reshape(3, 2))
    out, scale = s.sample(H.CoordsVelBox(np.zeros((2, 3)), np.zeros((2, 3)), np.eye(3)), replica_idx=2, state_idx=1, steps_done=5, n_steps=4)
    assert calls == ["x", "v", "box", ("params", [2.0, 3.0]), ("step", (2 << 40) + 5)] and scale is None
    assert out.velocities.shape == (2, 3)


def test_context_sampler_drives_the_water_sampler_like_the_reference():
    """fe/free_energy.py:1497-1528: before a replica is sampled the exchange mover gets the nonbonded parameters of the
    replica's state and the replica's step count; the proposals / acceptances of the call are recorded per state."""
    calls = []
    counts = {"p": 0, "a": 0}

    def run(n):
        counts["p"] += 1000
        counts["a"] += 7
        return np.zeros((1, 2, 3)), np.eye(3)[None]

    mover = types.SimpleNamespace(
        set_params=lambda p: calls.append(("water_params", np.asarray(p).tolist())), set_step=lambda s: calls.append(("water_step", s)),
        n_proposed=lambda: counts["p"], n_accepted=lambda: counts["a"],
    )
    baro = types.SimpleNamespace(set_step=lambda s: calls.append(("baro_step", s)), get_volume_scale_factor=lambda: 0.5)
    bp = types.SimpleNamespace(get_potential=lambda: "pot", set_params=lambda p: calls.append(("params", p.tolist())))
    intg = types.SimpleNamespace(set_step=lambda s: calls.append(("step", s)))
    ctx = types.SimpleNamespace(
        get_potentials=lambda: [bp], get_integrator=lambda: intg, get_barostat=lambda: baro, get_movers=lambda: [baro, mover],
        set_x_t=lambda x: None, set_v_t=lambda v: None, set_box=lambda b: None, multiple_steps=run, get_v_t=lambda: np.ones((2, 3)),
    )
    water = np.arange(3 * 2 * 4, dtype=float).reshape(3, 2, 4)
    s = H.ContextSampler(ctx, np.arange(6.0).reshape(3, 2), water_params_by_state=water)
    assert s.water_sampler is mover
    xvb = H.CoordsVelBox(np.zeros((2, 3)), np.zeros((2, 3)), np.eye(3))
    _, scale = s.sample(xvb, replica_idx=1, state_idx=2, steps_done=800, n_steps=400)
    assert scale == 0.5
    assert calls == [
        ("params", [4.0, 5.0]), ("step", (1 << 40) + 800), ("baro_step", 800), ("water_params", water[2].tolist()), ("water_step", 800),
    ]
    # the movers count frames like the reference (current_frame * steps_per_frame), the noise stream counts every step
    calls.clear()
    s.sample(xvb, replica_idx=1, state_idx=2, steps_done=1000, n_steps=400, mover_step=800)
    assert calls == [
        ("params", [4.0, 5.0]), ("step", (1 << 40) + 1000), ("baro_step", 800), ("water_params", water[2].tolist()), ("water_step", 800),
    ]
    counts["p"] -= 1000
    counts["a"] -= 7
    s.sample(xvb, replica_idx=0, state_idx=0, steps_done=800, n_steps=400)
    assert s.water_sampling_counts == {2: (7, 1000), 0: (7, 1000)}
    with pytest.raises(AssertionError, match="no exchange mover"):
        ctx2 = types.SimpleNamespace(**{**ctx.__dict__, "get_movers": lambda: [baro]})
        H.ContextSampler(ctx2, np.arange(6.0).reshape(3, 2), water_params_by_state=water)
