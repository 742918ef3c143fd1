"""CPU, world_size 2 over gloo: the N>1 host logic (energy-row all-gather + deterministic neighbour swaps) that
bench.py runs over NCCL on the GPU box."""

import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from timemachine_b200 import replica

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        states = np.arange(world)
        rng = np.random.default_rng(99)  # same seed on every rank
        history = []
        for it in range(20):
            mine = int(states[rank])
            cand = replica.candidate_states(mine, world)
            # synthetic energies: replica r prefers state (r + it) % world more and more
            energies = [float((k - ((rank + it) % world)) ** 2) * 3.0 for k in cand]
            row = replica.energy_row(world, cand, energies)
            u = replica.all_gather_rows(row, dist)
            assert u.shape == (world, world)
            np.testing.assert_array_equal(u[rank], row)
            states = replica.neighbour_swaps(u, states, 300.0, rng)
            history.append(states.copy())
        np.save(os.path.join(out_dir, f"hist_{rank}.npy"), np.array(history))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_agrees(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    h0 = np.load(tmp_path / "hist_0.npy")
    h1 = np.load(tmp_path / "hist_1.npy")
    np.testing.assert_array_equal(h0, h1)  # every rank reaches the same permutation with no extra traffic
    for states in h0:
        assert sorted(states.tolist()) == list(range(world))
    assert len({tuple(s) for s in h0}) > 1  # some swap was accepted


def test_neighbour_swaps_detailed_balance_direction():
    from timemachine_b200 import replica

    # replica 0 holds state 0 but is much lower in energy in state 1, and vice versa: the swap must be accepted
    u = np.array([[10.0, 0.0], [0.0, 10.0]])
    out = replica.neighbour_swaps(u, np.array([0, 1]), 300.0, np.random.default_rng(0))
    assert out.tolist() == [1, 0]
    # the reverse situation is (almost surely) rejected
    u = np.array([[0.0, 1000.0], [1000.0, 0.0]])
    out = replica.neighbour_swaps(u, np.array([0, 1]), 300.0, np.random.default_rng(0))
    assert out.tolist() == [0, 1]
    # non-evaluated entries (inf) never swap
    u = np.array([[0.0, np.inf], [np.inf, 0.0]])
    assert replica.neighbour_swaps(u, np.array([0, 1]), 300.0, np.random.default_rng(0)).tolist() == [0, 1]


def test_single_rank_paths():
    from timemachine_b200 import replica

    row = replica.energy_row(1, replica.candidate_states(0, 1), [1.5])
    assert replica.all_gather_rows(row).shape == (1, 1)
    assert replica.neighbour_swaps(row.reshape(1, 1), np.array([0]), 300.0, np.random.default_rng(0)).tolist() == [0]
    assert replica.i128_to_energy(1 << 36, 0) == 1.0
    assert replica.i128_to_energy((-(1 << 36)) & ((1 << 64) - 1), -1) == -1.0
    assert np.isnan(replica.i128_to_energy(0, 1))


def test_run_neighbor_swaps_matches_reference_golden():
    """replica.run_neighbor_swaps against the reference's own `_run_neighbor_swaps` (md/hrex.py:50-129), executed from the
    reference source by tests/golden/make_golden.py -> hrex.npz."""
    from pathlib import Path

    from timemachine_b200 import replica

    g = np.load(Path(__file__).parent / "golden" / "hrex.npz")
    final, proposed, accepted = replica.run_neighbor_swaps(
        g["start"], g["neighbor_pairs"], g["log_q"], g["pair_idxs"], g["uniform_samples"]
    )
    np.testing.assert_array_equal(final, g["final"])
    np.testing.assert_array_equal(proposed, g["proposed"])
    np.testing.assert_array_equal(accepted, g["accepted"])
    assert sorted(final.tolist()) == list(range(len(final)))  # still a permutation
    assert proposed.sum() == len(g["pair_idxs"]) and 0 < accepted.sum() < proposed.sum()


def test_native_neighbor_swap_loop_equals_the_python_statement():
    """`tmb_hrex_run_neighbor_swaps` (C ABI, host code) against `replica.run_neighbor_swaps` (the restatement of the
    reference's `_run_neighbor_swaps`, md/hrex.py:50-129, pinned by tests/golden/hrex_driver.npz): identical permutations
    and counters, including +inf (not evaluated), -inf and NaN entries of log_q."""
    from timemachine_b200 import replica as R

    rng = np.random.default_rng(5)
    for n_states, n_attempts in ((2, 8), (4, 64), (8, 512), (11, 1331)):
        pairs = [(s, s + 1) for s in range(n_states - 1)]
        if n_states == 2:
            pairs = [(0, 0), *pairs]
        for trial in range(6):
            log_q = rng.normal(0, 3.0, (n_states, n_states))
            if trial >= 2:
                log_q[rng.random(log_q.shape) < 0.3] = -np.inf  # energies of +inf: states that were not evaluated
            if trial >= 4:
                log_q[rng.random(log_q.shape) < 0.1] = np.nan
                log_q[rng.random(log_q.shape) < 0.1] = np.inf
            perm0 = rng.permutation(n_states)
            pidx = rng.integers(0, len(pairs), n_attempts)
            us = rng.random(n_attempts)
            with np.errstate(invalid="ignore"):
                a = R.run_neighbor_swaps(perm0, pairs, log_q, pidx, us)
            b = R.run_neighbor_swaps_native(perm0, pairs, log_q, pidx, us)
            for x, y in zip(a, b):
                np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
    with pytest.raises(RuntimeError, match="outside"):
        R.run_neighbor_swaps_native([0, 1], [(0, 5)], np.zeros((2, 2)), [0], [0.5])
