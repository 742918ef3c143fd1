"""Hamiltonian replica exchange over the one-replica-per-GPU layout.

Mirrors the reference's host driver - `timemachine/md/hrex.py` (HREX bookkeeping :132-240, diagnostics :243-394) and
`timemachine/fe/free_energy.py` (`compute_potential_matrix` :1148-1200, `verify_and_sanitize_potential_matrix`
:1203-1217, `run_sims_hrex` :1383-1618) and the on-disk frame store `timemachine/fe/stored_arrays.py` - with the same
names, argument meaning and return shapes, so a caller of the reference finds the same pieces.  What differs is where
the work runs:

  reference: ONE Context on ONE GPU; per frame every replica is loaded into it (set_x_t / set_v_t / set_box /
             set_params) and sampled in turn, then `execute_batch_sparse` fills U_kl (fe/free_energy.py:1485-1547).
  here:      replica k lives on rank k % world_size (its own Context, coordinates never leave the GPU's host); every
             rank samples its replicas, evaluates THEIR rows of U_kl (`execute_batch_sparse` on its own potential), one
             all-gather of the rows (K doubles per replica: NCCL on the GPU box, gloo in the CPU tests) gives every rank
             the full matrix, and every rank runs the same seeded batch of neighbour swaps (`replica.run_neighbor_swaps`,
             pinned to the reference's `_run_neighbor_swaps` by tests/golden/hrex.npz) - no further traffic.  A swap
             changes which PARAMETER SET a replica is sampled under.  world_size == 1 is the reference's sequential
             algorithm through the same code.

Frames are written by whichever rank sampled them into `<out_dir>/state_<s>/<iter>.npy`, the chunk layout of the
reference's StoredArrays (`get_chunk_path`: `<prefix>/<idx>.npy`), so the per-state trajectories are complete on a shared
filesystem without any frame crossing a collective.

The Langevin noise of this repository is counter-based (Philox keyed by atom and step, `LangevinIntegrator.set_step`):
replica k always draws from sub-stream k at its own step count, so trajectories do not depend on the layout - the same
seeds give bitwise the same frames on 1, 2 or 8 GPUs (tested).  The reference's cuRAND stream cannot offer that.

Swap proposals: the reference draws (pair_idxs, uniform_samples) with `jax.random` from `seed` (md/hrex.py:226-229); jax is
not part of this image, numpy's Philox generator seeded with the same integer is used instead (same distribution, not
the same stream).
"""

from __future__ import annotations

import tempfile
import time
import warnings
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Optional, Sequence

import numpy as np

from . import replica as _replica

BOLTZ = _replica.BOLTZ


class IndeterminateEnergyWarning(UserWarning):
    pass


@dataclass
class CoordsVelBox:
    coords: np.ndarray
    velocities: np.ndarray
    box: np.ndarray


# ---------------------------------------------------------------------------------------------------------------------
class StoredArrays:
    """Sequence of numpy arrays backed by one .npy file per chunk in a directory (reference fe/stored_arrays.py:14-100:
    same chunk naming, `<dir>/<idx>.npy`, so either implementation reads the other's directories)."""

    def __init__(self, path: Optional[Path] = None):
        self._tmp = None
        if path is None:
            self._tmp = tempfile.TemporaryDirectory()
            path = Path(self._tmp.name)
        self._dir = Path(path)
        self._dir.mkdir(parents=True, exist_ok=True)
        self._chunk_sizes: list[int] = []

    @staticmethod
    def get_chunk_path(path: Path, idx: int) -> Path:
        return (Path(path) / str(idx)).with_suffix(".npy")

    @classmethod
    def load(cls, path: Path) -> "StoredArrays":
        """Adopt the chunks 0, 1, 2, ... found in `path` (stops at the first missing index)."""
        sa = cls(path)
        idx = 0
        while cls.get_chunk_path(sa._dir, idx).exists():
            sa._chunk_sizes.append(len(np.load(cls.get_chunk_path(sa._dir, idx), mmap_mode="r")))
            idx += 1
        return sa

    def extend(self, xs) -> None:
        xs = np.asarray(xs)
        np.save(self.get_chunk_path(self._dir, len(self._chunk_sizes)), xs)
        self._chunk_sizes.append(len(xs))

    def __len__(self) -> int:
        return sum(self._chunk_sizes)

    def __iter__(self):
        for idx in range(len(self._chunk_sizes)):
            yield from np.load(self.get_chunk_path(self._dir, idx))

    def __getitem__(self, key):
        if isinstance(key, slice):
            raise NotImplementedError("slices are not implemented")
        if not isinstance(key, (int, np.integer)):
            raise ValueError("invalid subscript")
        key = range(len(self))[key]
        for idx, size in enumerate(self._chunk_sizes):
            if key < size:
                return np.load(self.get_chunk_path(self._dir, idx))[key]
            key -= size
        raise AssertionError("internal error")

    def __eq__(self, other) -> bool:
        return self._chunk_sizes == other._chunk_sizes and all(np.array_equal(a, b, equal_nan=True) for a, b in zip(self, other))


@dataclass
class Trajectory:
    """Frames of one state (reference fe/free_energy.py:265-291)."""

    frames: StoredArrays
    boxes: list
    final_velocities: Optional[np.ndarray] = None
    final_barostat_volume_scale_factor: Optional[float] = None


# ---------------------------------------------------------------------------------------------------------------------
class MonteCarloMove:
    """propose() -> (proposal, log acceptance probability); move() accepts with that probability (NumPy's global stream, like
    the reference's md/moves.py:52-85) and counts."""

    def __init__(self):
        self._n_proposed = 0
        self._n_accepted = 0

    def propose(self, x):
        raise NotImplementedError

    def move(self, x):
        proposal, log_acceptance_probability = self.propose(x)
        self._n_proposed += 1
        # NaN compares false: a proposal with an indeterminate weight is rejected
        if np.random.rand() < np.exp(log_acceptance_probability):
            self._n_accepted += 1
            return proposal
        return x

    def sample_chain(self, x, n_samples: int) -> list:
        out = []
        for _ in range(n_samples):
            x = self.move(x)
            out.append(x)
        return out

    @property
    def n_proposed(self) -> int:
        return self._n_proposed

    @property
    def n_accepted(self) -> int:
        return self._n_accepted

    @property
    def acceptance_fraction(self) -> float:
        return self._n_accepted / self._n_proposed if self._n_proposed else float("nan")


class MixtureOfMoves:
    """One move per step, chosen uniformly (md/moves.py:101-128)."""

    def __init__(self, moves: Sequence[MonteCarloMove]):
        self.moves = list(moves)

    @property
    def n_accepted_by_move(self) -> list:
        return [m.n_accepted for m in self.moves]

    @property
    def n_proposed_by_move(self) -> list:
        return [m.n_proposed for m in self.moves]

    def move(self, x):
        return self.moves[np.random.choice(len(self.moves))].move(x)

    def move_n(self, x, n: int):
        for idx in np.random.choice(len(self.moves), size=n, replace=True):
            x = self.moves[idx].move(x)
        return x


class NeighborSwapMove(MonteCarloMove):
    """Swap the replicas at a fixed pair of states (md/hrex.py:25-48).  The state of the move is the list replica-by-state."""

    def __init__(self, log_q: Callable, s_a: int, s_b: int):
        super().__init__()
        self.log_q = log_q
        self.s_a = s_a
        self.s_b = s_b

    def propose(self, state: list):
        s_a, s_b = self.s_a, self.s_b
        proposed = list(state)
        proposed[s_a], proposed[s_b] = state[s_b], state[s_a]
        r_a, r_b = state[s_a], state[s_b]
        log_q_diff = self.log_q(r_a, s_b) + self.log_q(r_b, s_a) - self.log_q(r_a, s_a) - self.log_q(r_b, s_b)
        return proposed, np.minimum(log_q_diff, 0.0)


@dataclass(frozen=True)
class HREX:
    """Replicas and the permutation state -> replica (reference md/hrex.py:132-240)."""

    replicas: list
    replica_idx_by_state: list

    @classmethod
    def from_replicas(cls, replicas: Sequence) -> "HREX":
        return HREX(list(replicas), list(range(len(replicas))))

    @property
    def state_replica_pairs(self) -> list:
        return [(s, self.replicas[r]) for s, r in enumerate(self.replica_idx_by_state)]

    def sample_replicas(self, sample_replica: Callable, replica_from_samples: Callable):
        """Sample every (state, replica) pair in state order; returns the updated HREX and the samples by state."""
        samples_by_state = [sample_replica(rep, s) for s, rep in self.state_replica_pairs]
        replicas = list(self.replicas)
        for s, samples in enumerate(samples_by_state):
            replicas[self.replica_idx_by_state[s]] = replica_from_samples(samples)
        return HREX(replicas, self.replica_idx_by_state), samples_by_state

    def attempt_neighbor_swaps(self, neighbor_pairs, log_q: Callable, n_swap_attempts: int):
        """The reference's plain version (md/hrex.py:155-188): a mixture of NeighborSwapMoves on log_q(replica_idx, state_idx),
        NumPy's global stream.  Same distribution as attempt_neighbor_swaps_fast, not the same sequence (as in the reference)."""
        move = MixtureOfMoves([NeighborSwapMove(log_q, s_a, s_b) for s_a, s_b in neighbor_pairs])
        replica_idx_by_state = move.move_n(list(self.replica_idx_by_state), n_swap_attempts)
        return HREX(self.replicas, replica_idx_by_state), list(zip(move.n_accepted_by_move, move.n_proposed_by_move))

    def attempt_neighbor_swaps_fast(self, neighbor_pairs, log_q_kl, n_swap_attempts: int, seed: int):
        """A batch of swap attempts, each between a uniformly chosen neighbour pair (md/hrex.py:195-235).
        Returns the updated HREX and [(accepted, proposed)] by pair."""
        rng = np.random.Generator(np.random.Philox(int(seed)))
        pair_idxs = rng.integers(0, len(neighbor_pairs), n_swap_attempts)
        uniform_samples = rng.random(n_swap_attempts)
        final, proposed, accepted = _replica.run_neighbor_swaps_native(
            np.asarray(self.replica_idx_by_state), np.asarray(neighbor_pairs), np.asarray(log_q_kl), pair_idxs, uniform_samples
        )
        return HREX(self.replicas, [int(r) for r in final]), list(zip(accepted.tolist(), proposed.tolist()))


def get_swap_attempts_per_iter_heuristic(n_states: int) -> int:
    return n_states**3  # md/hrex.py:386-394


def get_cumulative_replica_state_counts(replica_idx_by_state_by_iter) -> np.ndarray:
    """(iter, state, replica) -> how often `replica` has held `state` up to and including `iter` (md/hrex.py:272-286)."""
    a = np.asarray(replica_idx_by_state_by_iter)
    occupancy = a[:, :, None] == np.arange(a.shape[1])
    return np.cumsum(occupancy.astype(int), axis=0)


def estimate_transition_matrix(replica_idx_by_state_by_iter) -> np.ndarray:
    """(to state, from state) -> fraction of iterations in which the replica of `from` moved to `to` (md/hrex.py:289-306)."""
    a = np.asarray(replica_idx_by_state_by_iter)
    moved = a[:-1, None, :] == a[1:, :, None]
    return moved.sum(axis=0) / (a.shape[0] - 1)


def estimate_relaxation_time(transition_matrix: np.ndarray) -> float:
    """1 / (1 - second-largest eigenvalue of the symmetrised transition matrix) (md/hrex.py:309-334)."""
    assert np.allclose(transition_matrix.sum(axis=0), 1.0), "columns of transition matrix must sum to 1"
    mu = np.linalg.eigvalsh(0.5 * (transition_matrix + transition_matrix.T))
    return float(1.0 / (1.0 - mu[-2]))


def get_normalized_kl_divergence(replica_idx_by_state_by_iter) -> float:
    """Mean over states of KL(replica occupancy || uniform); 0 = every replica visits every state equally
    (md/hrex.py:243-269)."""
    counts = get_cumulative_replica_state_counts(replica_idx_by_state_by_iter)
    n_iters, n_states, _ = counts.shape
    p = counts[-1] / n_iters  # (state, replica)
    with np.errstate(divide="ignore", invalid="ignore"):
        plogp = np.where(p > 0, p * np.log(p), 0.0)
    col = p.sum(axis=0)  # scipy.stats.entropy normalises along axis 0
    ent = -(plogp / col).sum(axis=0) + np.log(col)
    return float(-np.mean(ent) + np.log(n_states))


def get_samples_by_iter_by_replica(samples_by_state_by_iter, replica_idx_by_state_by_iter):
    """(iter, state) -> samples  to  (replica, iter) -> samples (md/hrex.py:337-357)."""
    assert len(samples_by_state_by_iter) == len(replica_idx_by_state_by_iter)
    by_replica_by_iter = [
        [by_state[s] for s in np.argsort(perm)] for by_state, perm in zip(samples_by_state_by_iter, replica_idx_by_state_by_iter)
    ]
    return [list(xs) for xs in zip(*by_replica_by_iter)]


@dataclass
class WaterSamplingDiagnostics:
    """timemachine/md/exchange/exchange_mover.py:55-61: (accepted, proposed) water-exchange moves, [n_iters, n_states, 2]."""

    proposals_by_state_by_iter: np.ndarray

    @property
    def cumulative_proposals_by_state(self) -> np.ndarray:
        return np.sum(self.proposals_by_state_by_iter, axis=0)


@dataclass
class HREXDiagnostics:
    replica_idx_by_state_by_iter: list
    fraction_accepted_by_pair_by_iter: list
    # set when the sampler drives a water-exchange mover (SimulationResult.water_sampling_diagnostics, fe/free_energy.py:320,1615)
    water_sampling_diagnostics: Optional[WaterSamplingDiagnostics] = None

    @property
    def cumulative_swap_acceptance_rates(self) -> np.ndarray:
        n_accepted, n_proposed = np.moveaxis(np.array(self.fraction_accepted_by_pair_by_iter), -1, 0)
        return np.cumsum(n_accepted, axis=0) / np.cumsum(n_proposed, axis=0)

    @property
    def cumulative_replica_state_counts(self) -> np.ndarray:
        return get_cumulative_replica_state_counts(self.replica_idx_by_state_by_iter)

    @property
    def transition_matrix(self) -> np.ndarray:
        return estimate_transition_matrix(self.replica_idx_by_state_by_iter)

    @property
    def relaxation_time(self) -> float:
        return estimate_relaxation_time(self.transition_matrix)

    @property
    def normalized_kl_divergence(self) -> float:
        return get_normalized_kl_divergence(self.replica_idx_by_state_by_iter)


def _batches(n: int, batch_size: int):
    """timemachine/utils.py:6-13"""
    assert n >= 0 and batch_size > 0
    quot, rem = divmod(n, batch_size)
    for _ in range(quot):
        yield batch_size
    if rem:
        yield rem


def run_hrex(
    replicas: Sequence,
    sample_replica: Callable,
    replica_from_samples: Callable,
    neighbor_pairs: Sequence,
    get_log_q: Callable,
    n_samples: int,
    n_samples_per_iter: int,
    seed: int,
    n_swap_attempts_per_iter: Optional[int] = None,
):
    """The reference's generic HREX loop (md/hrex.py:397-491): per iteration a batch of neighbour swap attempts on the log
    weights of the current replicas (seed + iteration), then local sampling of every (state, replica) pair.

    sample_replica(replica, state_idx, n_samples) -> samples; replica_from_samples(samples) -> replica;
    get_log_q(replicas) -> (n_replicas, n_states) array, or a function (replica_idx, state_idx) -> log weight.
    Returns (samples by iteration and state, HREXDiagnostics)."""
    n_replicas = len(replicas)
    if n_swap_attempts_per_iter is None:
        n_swap_attempts_per_iter = get_swap_attempts_per_iter_heuristic(n_replicas)
    hrex = HREX.from_replicas(replicas)
    samples_by_state_by_iter: list = []
    replica_idx_by_state_by_iter: list = []
    fraction_accepted_by_pair_by_iter: list = []
    for iteration, n_samples_batch in enumerate(_batches(n_samples, n_samples_per_iter)):
        log_q = get_log_q(hrex.replicas)
        if callable(log_q):
            log_q_kl = np.array([[log_q(r, s) for s in range(n_replicas)] for r in range(n_replicas)])
        else:
            log_q_kl = np.asarray(log_q)
        hrex, fraction_accepted_by_pair = hrex.attempt_neighbor_swaps_fast(
            neighbor_pairs, log_q_kl, n_swap_attempts_per_iter, seed + iteration
        )
        hrex, samples_by_state = hrex.sample_replicas(
            lambda replica, state_idx: sample_replica(replica, state_idx, n_samples_batch), replica_from_samples
        )
        fraction_accepted_by_pair_by_iter.append(fraction_accepted_by_pair)
        replica_idx_by_state_by_iter.append(hrex.replica_idx_by_state)
        samples_by_state_by_iter.append(samples_by_state)
    return samples_by_state_by_iter, HREXDiagnostics(replica_idx_by_state_by_iter, fraction_accepted_by_pair_by_iter)


# ---------------------------------------------------------------------------------------------------------------------
def _sparse_batch_idxs(replica_idx_by_state, replica_idxs, max_delta_states: Optional[int]):
    """For each replica in `replica_idxs`: the states within max_delta_states of the state it currently holds (all states
    if None), as parallel (replica, state) index arrays - the bookkeeping of `compute_sparse`, fe/free_energy.py:1177-1192."""
    n_states = len(replica_idx_by_state)
    state_of_replica = np.argsort(replica_idx_by_state)
    rs, ss = [], []
    for r in replica_idxs:
        s0 = int(state_of_replica[r])
        lo, hi = (0, n_states - 1) if max_delta_states is None else (max(0, s0 - max_delta_states), min(n_states - 1, s0 + max_delta_states))
        for s in range(lo, hi + 1):
            rs.append(r)
            ss.append(s)
    return np.array(rs, dtype=np.uint32), np.array(ss, dtype=np.uint32)


def compute_potential_matrix(potential, hrex: HREX, params_by_state, max_delta_states: Optional[int] = None) -> np.ndarray:
    """(n_replicas, n_states) energies; element (k, l) is computed iff state l is within max_delta_states of replica k's
    current state, else +inf (reference fe/free_energy.py:1148-1200).  Single-process form: all replicas are local."""
    coords = np.array([xvb.coords for xvb in hrex.replicas])
    boxes = np.array([xvb.box for xvb in hrex.replicas])
    n = len(hrex.replicas)
    if max_delta_states is None:
        _, _, U_kl = potential.execute_batch(coords, params_by_state, boxes, False, False, True)
        return np.asarray(U_kl)
    cidx, pidx = _sparse_batch_idxs(hrex.replica_idx_by_state, range(n), max_delta_states)
    _, _, U = potential.execute_batch_sparse(coords, params_by_state, boxes, cidx, pidx, False, False, True)
    U_kl = np.full((n, n), np.inf)
    U_kl[cidx, pidx] = U
    return U_kl


def verify_and_sanitize_potential_matrix(U_kl, replica_idx_by_state, abs_energy_threshold: float = 1e9) -> np.ndarray:
    """Replicas must have finite, sane energies in the state they hold; NaN elsewhere becomes +inf
    (reference fe/free_energy.py:1203-1217)."""
    replica_energies = np.diagonal(U_kl[np.asarray(replica_idx_by_state)])
    assert np.all(np.isfinite(replica_energies)), "Replicas have non-finite energies"
    assert np.all(np.abs(replica_energies) < abs_energy_threshold), "Energies larger in magnitude than tolerated"
    if np.any(np.isnan(U_kl)):
        warnings.warn("Encountered NaNs in potential matrix. Replacing each instance with inf", IndeterminateEnergyWarning)
        U_kl = np.where(np.isnan(U_kl), np.inf, U_kl)
    return U_kl


# ---------------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class HREXMDParams:
    """The fields of the reference's MDParams / HREXParams that `run_sims_hrex` reads (fe/free_energy.py:70-175)."""

    n_frames: int
    steps_per_frame: int
    n_eq_steps: int = 0
    seed: int = 2024
    max_delta_states: Optional[int] = 4
    n_swap_attempts_per_iter: Optional[int] = None


class ContextSampler:
    """Drives one custom_ops.Context the way `sample_replica` does in the reference (fe/free_energy.py:1485-1540): load
    the replica, bind the parameters of the state it holds, run MD, read the last frame back.  One per rank."""

    SUBSTREAM_BITS = 40  # noise counter = (replica << 40) + steps taken by that replica: 2^40 steps per replica

    def __init__(self, context, params_by_state, water_params_by_state=None):
        self.context = context
        bps = context.get_potentials()
        assert len(bps) == 1, "HREX expects the whole system in one SummedPotential (fe/free_energy.py:1436-1438)"
        self.bound = bps[0]
        self.potential = self.bound.get_potential()
        self.params_by_state = np.ascontiguousarray(params_by_state, dtype=np.float64)
        self.integrator = context.get_integrator()
        self.barostat = context.get_barostat()
        # water sampling (fe/free_energy.py:1448-1466, 1497-1528): the exchange mover among the context's movers gets the
        # nonbonded parameters of the state a replica is sampled under, [n_states, N, 4] (get_water_sampler_params)
        movers = context.get_movers() if hasattr(context, "get_movers") else []
        self.water_sampler = next((m for m in movers if hasattr(m, "n_proposed") and hasattr(m, "set_params")), None)
        self.water_params_by_state = None if water_params_by_state is None else np.ascontiguousarray(water_params_by_state, dtype=np.float64)
        if self.water_params_by_state is not None:
            assert self.water_sampler is not None, "water parameters given but the context has no exchange mover"
            assert len(self.water_params_by_state) == len(self.params_by_state)
        self.water_sampling_counts = {}  # state -> (accepted, proposed) during the most recent sample() under that state

    def sample(self, xvb: CoordsVelBox, replica_idx: int, state_idx: int, steps_done: int, n_steps: int, mover_step: Optional[int] = None):
        """steps_done positions the integrator's counter-based noise stream (it includes the equilibration steps of frame
        0); mover_step is what the barostat and the water sampler are told, `current_frame * steps_per_frame` in the
        reference (fe/free_energy.py:1497-1510: equilibration steps are NOT counted, so the phase of the movers relative
        to the frames matches the reference for any n_eq_steps).  Defaults to steps_done."""
        if mover_step is None:
            mover_step = steps_done
        ctx = self.context
        ctx.set_x_t(xvb.coords)
        ctx.set_v_t(xvb.velocities)
        ctx.set_box(xvb.box)
        self.bound.set_params(self.params_by_state[state_idx])
        self.integrator.set_step((replica_idx << self.SUBSTREAM_BITS) + steps_done)
        if self.barostat is not None:
            self.barostat.set_step(mover_step)
        if self.water_sampler is not None:
            if self.water_params_by_state is not None:
                self.water_sampler.set_params(self.water_params_by_state[state_idx])
            self.water_sampler.set_step(mover_step)
            proposed0, accepted0 = self.water_sampler.n_proposed(), self.water_sampler.n_accepted()
        xs, boxes = ctx.multiple_steps(n_steps)
        if self.water_sampler is not None:
            self.water_sampling_counts[state_idx] = (
                self.water_sampler.n_accepted() - accepted0, self.water_sampler.n_proposed() - proposed0,
            )
        scale = self.barostat.get_volume_scale_factor() if self.barostat is not None else None
        return CoordsVelBox(xs[-1], ctx.get_v_t(), boxes[-1]), scale

    def energies(self, xvbs: Sequence[CoordsVelBox], coords_batch_idxs, params_batch_idxs) -> np.ndarray:
        coords = np.array([x.coords for x in xvbs])
        boxes = np.array([x.box for x in xvbs])
        _, _, U = self.potential.execute_batch_sparse(
            coords, self.params_by_state, boxes, np.asarray(coords_batch_idxs, np.uint32), np.asarray(params_batch_idxs, np.uint32), False, False, True
        )
        return np.asarray(U)


class ResidentCoordsVelBox:
    """Stands for a replica whose coordinates, velocities and box live in a Context's device buffers.  `coords`,
    `velocities` and `box` read them back on demand; once the sampler loads another replica into the context the values are
    captured to the host first (`_materialize`), so the object stays valid."""

    def __init__(self, sampler, replica_idx):
        self._sampler = sampler
        self.replica_idx = replica_idx
        self._host = None

    def _materialize(self):
        if self._host is None:
            ctx = self._sampler.context
            self._host = CoordsVelBox(np.asarray(ctx.get_x_t()), np.asarray(ctx.get_v_t()), np.asarray(ctx.get_box()))
        return self._host

    @property
    def coords(self):
        return self._host.coords if self._host is not None else np.asarray(self._sampler.context.get_x_t())

    @property
    def velocities(self):
        return self._host.velocities if self._host is not None else np.asarray(self._sampler.context.get_v_t())

    @property
    def box(self):
        return self._host.box if self._host is not None else np.asarray(self._sampler.context.get_box())


class DeviceResidentSampler(ContextSampler):
    """ContextSampler for the one-replica-per-GPU layout: the replica a rank owns never leaves the device.

    The reference reloads every replica into its single Context each frame (set_x_t / set_v_t / set_box / set_params,
    fe/free_energy.py:1485-1496) because all replicas share one GPU.  With a GPU per replica none of those copies is
    needed: the state stays in the Context's buffers, all K parameter sets are resident in HBM ([K, P] doubles) and a swap
    is a device-to-device re-bind (`BoundPotential.set_params_device`, reference bound_potential.cu:139-147); the energies
    of U_kl are energy-only evaluations on the resident coordinates (`Potential.execute_device` with du_dx = du_dp = null)
    read back as one small D2H copy.  Everything is enqueued on one stream.  Falls back to loading from the host when a rank
    owns several replicas (the previously resident one is captured to the host first)."""

    def __init__(self, context, params_by_state, device, stream=None, water_params_by_state=None):
        super().__init__(context, params_by_state, water_params_by_state)
        import torch

        self._torch = torch
        self.device = torch.device(device)
        self.stream = stream if stream is not None else torch.cuda.Stream(device=self.device)
        context.set_stream(self.stream.cuda_stream)
        self.d_x, self.d_v, self.d_box = context.device_state()
        self.d_params = torch.from_numpy(self.params_by_state).to(self.device)  # [K, P]
        self.n_atoms = None
        self.n_params = int(self.params_by_state.shape[1])
        self.d_u = torch.zeros(2 * max(4, 2 * len(self.params_by_state)), dtype=torch.int64, device=self.device)  # int128 slots
        self._resident = None  # the ResidentCoordsVelBox whose replica is in the context

    def sample(self, xvb, replica_idx: int, state_idx: int, steps_done: int, n_steps: int, mover_step: Optional[int] = None):
        if mover_step is None:
            mover_step = steps_done
        ctx = self.context
        if not (xvb is self._resident and self._resident._host is None):
            # not what the context holds right now: keep the resident replica's state on the host, load this one
            if self._resident is not None:
                self._resident._materialize()
            x, v, b = xvb.coords, xvb.velocities, xvb.box
            ctx.set_x_t(x)
            ctx.set_v_t(v)
            ctx.set_box(b)
            self.n_atoms = int(np.asarray(x).shape[0])
        self.bound.set_params_device(self.d_params[state_idx].data_ptr(), self.n_params, self.stream.cuda_stream)
        self.integrator.set_step((replica_idx << self.SUBSTREAM_BITS) + steps_done)
        if self.barostat is not None:
            self.barostat.set_step(mover_step)
        if self.water_sampler is not None:
            if self.water_params_by_state is not None:
                self.water_sampler.set_params(self.water_params_by_state[state_idx])
            self.water_sampler.set_step(mover_step)
            proposed0, accepted0 = self.water_sampler.n_proposed(), self.water_sampler.n_accepted()
        ctx.multiple_steps(n_steps, n_steps + 1)  # nothing is copied out
        if self.water_sampler is not None:
            self.water_sampling_counts[state_idx] = (
                self.water_sampler.n_accepted() - accepted0, self.water_sampler.n_proposed() - proposed0,
            )
        scale = self.barostat.get_volume_scale_factor() if self.barostat is not None else None
        self._resident = ResidentCoordsVelBox(self, replica_idx)
        return self._resident, scale

    def energies(self, xvbs, coords_batch_idxs, params_batch_idxs) -> np.ndarray:
        torch = self._torch
        out = np.empty(len(coords_batch_idxs))
        resident_pairs = []
        for slot, (c, p) in enumerate(zip(coords_batch_idxs, params_batch_idxs)):
            if xvbs[int(c)] is self._resident and self._resident._host is None:
                resident_pairs.append((slot, int(p)))
        if len(self.d_u) < 2 * len(resident_pairs):
            self.d_u = torch.zeros(2 * len(resident_pairs), dtype=torch.int64, device=self.device)
        for n, (slot, p) in enumerate(resident_pairs):
            self.potential.execute_device(
                self.n_atoms, self.n_params, self.d_x, self.d_params[p].data_ptr(), self.d_box, 0, 0, self.d_u.data_ptr() + 16 * n,
                self.stream.cuda_stream,
            )
        if resident_pairs:
            with torch.cuda.stream(self.stream):
                host = self.d_u[: 2 * len(resident_pairs)].cpu().numpy()  # the frame's only device -> host copy
            for n, (slot, _) in enumerate(resident_pairs):
                out[slot] = _replica.i128_to_energy(int(host[2 * n]), int(host[2 * n + 1]))
        rest = [i for i in range(len(out)) if i not in {s for s, _ in resident_pairs}]
        if rest:  # replicas that are not in the context right now: the host path
            out[rest] = super().energies(xvbs, np.asarray(coords_batch_idxs)[rest], np.asarray(params_batch_idxs)[rest])
        return out


def _all_gather_rows(local_rows: np.ndarray, owners: list, n_replicas: int, dist, device) -> np.ndarray:
    """local_rows[i] is the row of replica owners[rank][i].  Every rank contributes a fixed-size block (ranks with fewer
    replicas pad with +inf rows), one all-gather, and the rows are put back in replica order."""
    world = 1 if dist is None or not dist.is_initialized() else dist.get_world_size()
    if world == 1:
        return local_rows
    import torch

    per_rank = max(len(o) for o in owners)
    block = np.full((per_rank, n_replicas), np.inf)
    block[: len(local_rows)] = local_rows
    t = torch.from_numpy(block.reshape(-1))
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * block.size, dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, t)
    gathered = out.cpu().numpy().reshape(world, per_rank, n_replicas)
    U = np.full((n_replicas, n_replicas), np.inf)
    for r, mine in enumerate(owners):
        for i, k in enumerate(mine):
            U[k] = gathered[r, i]
    return U


def run_sims_hrex(
    sampler,
    replicas: Sequence[CoordsVelBox],
    temperature: float,
    md_params: HREXMDParams,
    out_dir: Optional[Path] = None,
    dist=None,
    device=None,
    print_diagnostics_interval: Optional[int] = None,
    on_iteration: Optional[Callable] = None,
    store_frames: bool = True,
    replica_idx_by_state: Optional[Sequence[int]] = None,
):
    """Nearest-neighbour HREX over len(replicas) states (reference run_sims_hrex, fe/free_energy.py:1383-1618, up to and
    including the diagnostics; the BAR analysis that follows it there is host science outside this path).

    sampler: this rank's ContextSampler (or anything with its `sample` / `energies` methods); `replicas` holds the initial
    (coords, velocities, box) of every state - only the ones this rank owns (k % world_size == rank) are used.
    on_iteration(frame, U_kl, hrex) is called on every rank after the energies of an iteration are known, before its swaps.
    Returns ([Trajectory by state], HREXDiagnostics, final HREX).  Trajectories are backed by `<out_dir>/state_<s>`; on
    ranks > 0 only the frames those ranks sampled exist until every rank has finished (shared filesystem), which the
    final barrier guarantees.  store_frames=False keeps nothing on disk and returns None for the trajectories (throughput
    runs with a DeviceResidentSampler: no frame ever leaves the GPU); the diagnostics and the final HREX are as usual.
    replica_idx_by_state continues an earlier run's permutation (default: replica k starts in state k).
    """
    n_states = len(replicas)
    world = 1 if dist is None or not dist.is_initialized() else dist.get_world_size()
    rank = 0 if world == 1 else dist.get_rank()
    owners = [[k for k in range(n_states) if k % world == r] for r in range(world)]
    mine = owners[rank]
    n_swaps = md_params.n_swap_attempts_per_iter or get_swap_attempts_per_iter_heuristic(n_states)
    neighbor_pairs = [(s, s + 1) for s in range(n_states - 1)]
    if n_states == 2:
        neighbor_pairs = [(0, 0), *neighbor_pairs]  # identity move for aperiodicity (fe/free_energy.py:1455-1457)

    tmp = None
    if out_dir is None and store_frames:
        assert world == 1, "ranks must share an out_dir"
        tmp = tempfile.TemporaryDirectory()
        out_dir = Path(tmp.name)
    out_dir = Path(out_dir) if store_frames else Path("/nonexistent")
    state_dirs = [out_dir / f"state_{s}" for s in range(n_states)]
    for d in state_dirs if store_frames else []:
        d.mkdir(parents=True, exist_ok=True)

    hrex = HREX.from_replicas([replicas[k] if k in mine else None for k in range(n_states)])
    if replica_idx_by_state is not None:
        assert sorted(replica_idx_by_state) == list(range(n_states)), "replica_idx_by_state must be a permutation"
        hrex = HREX(hrex.replicas, [int(r) for r in replica_idx_by_state])
    steps_done = {k: 0 for k in mine}
    box_dirs = [out_dir / f"state_{s}_boxes" for s in range(n_states)]
    for d in box_dirs if store_frames else []:
        d.mkdir(parents=True, exist_ok=True)
    water_dirs = [out_dir / f"state_{s}_water_sampling" for s in range(n_states)]  # created on first use
    replica_idx_by_state_by_iter, fraction_accepted_by_pair_by_iter = [], []
    kT = BOLTZ * temperature
    t_begin = t_last = time.perf_counter()

    import inspect

    takes_mover_step = "mover_step" in inspect.signature(sampler.sample).parameters
    for frame in range(md_params.n_frames):
        state_of_replica = np.argsort(hrex.replica_idx_by_state)
        local = list(hrex.replicas)
        for k in mine:
            s = int(state_of_replica[k])
            n_steps = md_params.steps_per_frame + (md_params.n_eq_steps if frame == 0 else 0)
            extra = {"mover_step": frame * md_params.steps_per_frame} if takes_mover_step else {}
            xvb, scale = sampler.sample(local[k], k, s, steps_done[k], n_steps, **extra)
            steps_done[k] += n_steps
            local[k] = xvb
            if not store_frames:
                continue
            # one chunk per (state, iteration), written by whoever sampled it: the reference's StoredArrays layout
            np.save(StoredArrays.get_chunk_path(state_dirs[s], frame), xvb.coords[None])
            np.save(StoredArrays.get_chunk_path(box_dirs[s], frame), xvb.box[None])
            water_counts = getattr(sampler, "water_sampling_counts", None)
            if water_counts is not None and s in water_counts:  # (accepted, proposed) of this call (fe/free_energy.py:1521-1528)
                water_dirs[s].mkdir(parents=True, exist_ok=True)
                np.save(StoredArrays.get_chunk_path(water_dirs[s], frame), np.asarray(water_counts[s], dtype=np.int32)[None])
            if frame == md_params.n_frames - 1:
                np.savez(out_dir / f"final_state_{s}.npz", velocities=xvb.velocities, scale=np.nan if scale is None else scale)
        hrex = HREX(local, hrex.replica_idx_by_state)

        # rows of U_kl for the replicas that live here, then one all-gather
        cidx, pidx = _sparse_batch_idxs(hrex.replica_idx_by_state, mine, md_params.max_delta_states)
        local_of = {k: i for i, k in enumerate(mine)}
        lidx = np.array([local_of[int(c)] for c in cidx], dtype=np.uint32)
        rows = np.full((len(mine), n_states), np.inf)
        if len(mine):
            rows[lidx, pidx] = sampler.energies([local[k] for k in mine], lidx, pidx)
        U_kl_raw = _all_gather_rows(rows, owners, n_states, dist, device)
        U_kl = verify_and_sanitize_potential_matrix(U_kl_raw, hrex.replica_idx_by_state)
        log_q_kl = -U_kl / kT
        if on_iteration is not None:
            on_iteration(frame, U_kl, hrex)

        replica_idx_by_state_by_iter.append(list(hrex.replica_idx_by_state))
        if n_states == 1:
            fraction = []  # a single window: nothing to swap with
        else:
            hrex, fraction = hrex.attempt_neighbor_swaps_fast(neighbor_pairs, log_q_kl, n_swaps, md_params.seed + frame + 1)
        if n_states == 2:
            fraction = fraction[1:]
        fraction_accepted_by_pair_by_iter.append(fraction)

        if print_diagnostics_interval and rank == 0 and (frame + 1) % print_diagnostics_interval == 0:
            now = time.perf_counter()
            acc = np.sum(fraction_accepted_by_pair_by_iter, axis=0)
            rates = " |".join(f"{100.0 * a / p:5.1f}%" if p else "  nan" for a, p in acc)
            print(f"Frame {frame + 1}: {(now - t_begin) / (frame + 1):.2f} s/frame ({(now - t_last) / print_diagnostics_interval:.2f} since last)")
            print("HREX acceptance rates, average:", rates)
            print("HREX replica permutation      :", hrex.replica_idx_by_state)
            t_last = now

    if world > 1:
        dist.barrier()  # every chunk is on disk
    diagnostics = HREXDiagnostics(replica_idx_by_state_by_iter, fraction_accepted_by_pair_by_iter)
    if not store_frames:
        return None, diagnostics, hrex
    trajectories = []
    for s in range(n_states):
        frames = StoredArrays.load(state_dirs[s])
        boxes = list(StoredArrays.load(box_dirs[s]))
        final = np.load(out_dir / f"final_state_{s}.npz")
        scale = float(final["scale"])
        trajectories.append(Trajectory(frames, boxes, final["velocities"], None if np.isnan(scale) else scale))
    if all(d.is_dir() for d in water_dirs):
        per_state = [np.array(list(StoredArrays.load(d)), dtype=np.int32) for d in water_dirs]  # [n_iters, 2] each
        diagnostics.water_sampling_diagnostics = WaterSamplingDiagnostics(np.stack(per_state, axis=1))
    if tmp is not None:
        for t in trajectories:
            t.frames._tmp = tmp  # keep the temporary directory alive as long as any trajectory is
    return trajectories, diagnostics, hrex
