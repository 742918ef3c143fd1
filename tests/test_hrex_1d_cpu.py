"""CPU: the generic HREX loop `run_hrex` on one-dimensional targets, modelled on the reference's tests/hrex/test_hrex_1d.py:
every state's samples follow its own distribution (KS test on thinned chains), swap acceptance is what the overlap implies,
replicas spend their time evenly over the states, NaN log weights are never accepted."""

from dataclasses import dataclass, replace

import numpy as np
import pytest
import scipy.stats
from scipy.special import logsumexp

from timemachine_b200.hrex import run_hrex


@dataclass
class GaussianMixture:
    locs: np.ndarray
    scales: np.ndarray
    log_weights: np.ndarray

    def sample(self, rng, n):
        probs = np.exp(self.log_weights - logsumexp(self.log_weights))
        comp = rng.choice(len(self.locs), p=probs, size=n)
        return rng.normal(self.locs, self.scales, size=(n, len(self.locs)))[np.arange(n), comp]

    def log_q(self, x):
        lq = -((np.atleast_1d(x)[:, None] - self.locs) ** 2) / (2 * self.scales**2)
        return logsumexp(lq + self.log_weights, axis=1).item()


def gaussian(loc, scale, log_weight=0.0):
    return GaussianMixture(np.array([loc]), np.array([scale]), np.array([log_weight]))


def run(states, initial, seed, radius=0.1, n_samples=6000, per_iter=20, poison=False):
    rng = np.random.default_rng(seed)
    idx = list(range(len(states)))
    # (0, 0) keeps a fixed number of neighbour swaps aperiodic when acceptance approaches 100 % (test_hrex_1d.py:100-102)
    pairs = [(0, 0), *zip(idx, idx[1:])]

    def log_q_of(state, x):
        return float("nan") if poison else state.log_q(x)

    def sample_replica(x, state_idx, n):
        target, out = states[state_idx], []
        for _ in range(n):  # Metropolis chain with a Gaussian proposal
            xp = x + radius * rng.normal()
            if np.log(rng.random()) < target.log_q(xp) - target.log_q(x):
                x = xp
            out.append(x)
        return out

    def get_log_q(replicas):
        return np.array([[log_q_of(states[s], replicas[r]) for s in idx] for r in idx])

    samples, diag = run_hrex(initial, sample_replica, lambda xs: xs[-1], pairs, get_log_q, n_samples, per_iter, seed)
    diag = replace(diag, fraction_accepted_by_pair_by_iter=[f[1:] for f in diag.fraction_accepted_by_pair_by_iter])
    return np.concatenate(samples, axis=1), diag, rng


@pytest.mark.parametrize("seed", range(3))
def test_different_distributions_same_free_energy(seed):
    locs = [0.0, 0.5, 1.0]
    states = [gaussian(loc, 0.3) for loc in locs]
    samples_by_state, diag, rng = run(states, locs, seed)
    assert samples_by_state.shape == (3, 6000)
    tau = round(1 / 0.1**2)
    p = [scipy.stats.ks_2samp(x[tau::tau], s.sample(rng, 6000)).pvalue for x, s in zip(samples_by_state, states)]
    np.testing.assert_array_less(0.002, p)
    final = diag.cumulative_swap_acceptance_rates[-1]
    np.testing.assert_array_less(0.17, final)  # ~0.2 - 0.22 (the reference asserts 0.2 on 10 000 samples of its own streams)
    np.testing.assert_array_less(np.abs(final - final.mean()), 0.03)
    density = diag.cumulative_replica_state_counts[-1] / diag.cumulative_replica_state_counts.shape[0]
    np.testing.assert_array_less(np.abs(density - density.mean()), 0.25)
    assert diag.relaxation_time > 0 and 0 <= diag.normalized_kl_divergence < 0.5


def test_same_distributions_different_free_energies():
    states = [gaussian(0.0, 0.3, w) for w in (-1.0, 0.0, 1.0)]
    _, diag, _ = run(states, [0.0] * 3, 1, n_samples=2000)
    assert np.all(diag.cumulative_swap_acceptance_rates == 1.0)  # the difference of log q over a swap is always zero


def test_mixture_is_crossed_through_the_broad_state():
    """Two narrow modes with no overlap: local moves alone never leave the first, exchange with a broad state finds both."""
    states = [GaussianMixture(np.array([0.0, 1.0]), np.array([0.1, 0.1]), np.zeros(2)), gaussian(0.5, 0.5)]
    samples_by_state, _, _ = run(states, [0.0, 0.0], 3, n_samples=12000)
    frac_right = np.mean(samples_by_state[0] > 0.5)
    assert 0.25 < frac_right < 0.75, frac_right


def test_nan_log_weights_are_never_accepted():
    states = [gaussian(loc, 0.3) for loc in (0.0, 0.5, 1.0)]
    _, diag, _ = run(states, [0.0, 0.5, 1.0], 0, n_samples=400, poison=True)
    np.testing.assert_array_equal(diag.cumulative_swap_acceptance_rates[-1], 0.0)


def test_plain_and_fast_swap_batches_agree_statistically():
    """HREX.attempt_neighbor_swaps (a mixture of NeighborSwapMoves, md/hrex.py:155-188) against attempt_neighbor_swaps_fast on
    the same log weights: same acceptance rate per pair, same distribution of final permutations."""
    from timemachine_b200.hrex import HREX, NeighborSwapMove

    rng = np.random.default_rng(0)
    n = 4
    log_q_kl = rng.normal(0, 1.0, (n, n))
    pairs = [(0, 1), (1, 2), (2, 3)]
    hrex = HREX.from_replicas(list("abcd"))
    np.random.seed(1)
    acc_plain, acc_fast = np.zeros((3, 2)), np.zeros((3, 2))
    perms_plain, perms_fast = {}, {}
    for it in range(300):
        h1, f1 = hrex.attempt_neighbor_swaps(pairs, lambda r, s: log_q_kl[r, s], 40)
        h2, f2 = hrex.attempt_neighbor_swaps_fast(pairs, log_q_kl, 40, seed=it)
        acc_plain += np.array(f1)
        acc_fast += np.array(f2)
        perms_plain[tuple(h1.replica_idx_by_state)] = perms_plain.get(tuple(h1.replica_idx_by_state), 0) + 1
        perms_fast[tuple(h2.replica_idx_by_state)] = perms_fast.get(tuple(h2.replica_idx_by_state), 0) + 1
        assert sorted(h1.replica_idx_by_state) == [0, 1, 2, 3] and h1.replicas == hrex.replicas
    np.testing.assert_allclose(acc_plain[:, 0] / acc_plain[:, 1], acc_fast[:, 0] / acc_fast[:, 1], atol=0.03)
    assert abs(acc_plain[:, 1].sum() - 300 * 40) < 1e-9
    top = max(perms_fast, key=perms_fast.get)
    assert abs(perms_plain.get(top, 0) - perms_fast[top]) < 60
    # a single move: the acceptance probability is min(1, exp(sum of the swapped log weights - sum of the current ones))
    move = NeighborSwapMove(lambda r, s: log_q_kl[r, s], 0, 1)
    proposed, log_p = move.propose([0, 1, 2, 3])
    assert proposed == [1, 0, 2, 3]
    assert log_p == pytest.approx(min(0.0, log_q_kl[0, 1] + log_q_kl[1, 0] - log_q_kl[0, 0] - log_q_kl[1, 1]))
