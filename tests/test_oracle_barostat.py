"""CPU checks of the barostat oracle (oracle/tm_oracle.py) against the golden vectors produced by
the reference's own Python CentroidRescaler (tests/golden/make_golden.py -> barostat.npz) and against the closed forms
of the acceptance rule.  The GPU parity tests (tests/test_barostat_gpu.py) then hold the CUDA barostat to this oracle
and, bit for bit, to the compiled reference."""

from pathlib import Path

import numpy as np
import pytest

from oracle import tm_oracle as O

GOLDEN = Path(__file__).parent / "golden" / "barostat.npz"


@pytest.fixture(scope="module")
def golden():
    g = dict(np.load(GOLDEN))
    offsets = np.concatenate([[0], np.cumsum(g["group_sizes"])])
    g["groups"] = [np.arange(offsets[i], offsets[i + 1]) for i in range(len(g["group_sizes"]))]
    return g


def test_scale_centroids_matches_reference_python(golden):
    out = O.barostat_scale_centroids(golden["coords"], golden["groups"], golden["center"], float(golden["scale"]))
    np.testing.assert_allclose(out, golden["scaled"], rtol=0, atol=1e-12)
    cents = np.array([golden["coords"][g].mean(axis=0) for g in golden["groups"]])
    np.testing.assert_allclose(cents, golden["centroids"], rtol=0, atol=1e-12)


def test_proposal_is_centroid_scaling_plus_home_box_imaging(golden):
    coords, groups = golden["coords"], golden["groups"]
    box = np.eye(3) * 3.1
    volume_scale, rand0 = 0.9, 0.8131
    x_prop, box_prop, volume, delta, used = O.barostat_propose(coords, box, groups, volume_scale, rand0, adaptive=False)
    assert used == volume_scale
    np.testing.assert_allclose(volume, 3.1**3, rtol=1e-6)
    np.testing.assert_allclose(delta, volume_scale * 2 * (rand0 - 0.5), rtol=1e-6)
    scale = np.cbrt((volume + delta) / volume)
    np.testing.assert_allclose(np.diag(box_prop), 3.1 * scale, rtol=1e-6)
    assert box_prop[0, 1] == 0 and box_prop[2, 0] == 0
    # float64 statement of the same move, built on the function that is pinned to the reference above
    expect = O.barostat_scale_centroids(coords, groups, np.full(3, 3.1 / 2), scale)
    for g in groups:
        centroid = expect[g].mean(axis=0)
        expect[g] -= np.diag(box_prop) * np.floor(centroid / np.diag(box_prop))
    np.testing.assert_allclose(x_prop, expect, rtol=0, atol=2e-6)
    # every moved centroid lies in the scaled home box, molecules are moved rigidly
    for g in groups:
        c = x_prop[g].mean(axis=0)
        assert np.all(c > -1e-5) and np.all(c < np.diag(box_prop) + 1e-5)
        d0 = coords[g][:, None, :] - coords[g][None, :, :]
        d1 = x_prop[g][:, None, :] - x_prop[g][None, :, :]
        np.testing.assert_allclose(d1, d0, rtol=0, atol=1e-12)


def test_proposal_default_volume_scale_and_ungrouped_atoms(golden):
    coords, groups = golden["coords"], golden["groups"]
    box = np.eye(3) * 3.1
    x_prop, _, volume, delta, used = O.barostat_propose(coords, box, groups[:-5], 0.0, 0.25, adaptive=True)
    np.testing.assert_allclose(used, 0.01 * float(volume), rtol=1e-12)  # "0.0 means 1 % of the volume" (k_barostat.cuh:107-109)
    np.testing.assert_allclose(delta, used * 2 * (0.25 - 0.5), rtol=1e-6)
    left_out = np.concatenate(groups[-5:])
    np.testing.assert_array_equal(x_prop[left_out], coords[left_out])
    # rand0 == 0.5: no volume change, only the imaging into the home box acts
    x_same, box_same, _, delta0, _ = O.barostat_propose(coords, box, groups, 1.0, 0.5, adaptive=False)
    assert delta0 == 0 and np.array_equal(box_same, box)
    shift = x_same - coords
    np.testing.assert_allclose(shift, np.round(shift / 3.1) * 3.1, rtol=0, atol=1e-5)


def test_acceptance_rule_closed_form():
    volume, n_mols, temperature, pressure = np.float32(27.0), 100, 300.0, 1.0
    kt = O.BOLTZ * temperature
    p = pressure * O.AVOGADRO * 1e-25
    rng = np.random.default_rng(5)
    for _ in range(200):
        delta = np.float32(rng.uniform(-0.5, 0.5))
        du = rng.uniform(-30, 30)
        u0 = int(rng.integers(-(1 << 50), 1 << 50))
        u1 = u0 + int(round(du * 2**36))
        rand1 = rng.random()
        accepted, w = O.barostat_accepts(u0, u1, volume, delta, n_mols, temperature, pressure, rand1)
        w64 = du + p * float(delta) - n_mols * kt * np.log((float(volume) + float(delta)) / float(volume))
        np.testing.assert_allclose(w, w64, rtol=2e-5, atol=2e-4)
        if abs(w64) > 1e-3 and abs(np.float32(rand1) - np.exp(-w64 / kt)) > 1e-3:
            assert accepted == (not (w64 > 0 and rand1 > np.exp(-w64 / kt)))
    # downhill moves are always accepted; an overflowed energy rejects unless the Metropolis draw is exactly zero
    assert O.barostat_accepts(0, -(1 << 40), volume, np.float32(0.1), n_mols, temperature, pressure, 0.999999)[0]
    overflowed = (1 << 63) - 1
    accepted, w = O.barostat_accepts(0, overflowed, volume, np.float32(0.1), n_mols, temperature, pressure, 1e-30)
    assert not accepted and np.isinf(w)


def test_adaptive_volume_scale_rule():
    assert O.barostat_adapt(1.0, 9, 0, 27.0) == (1.0, 9, 0)  # fewer than 10 attempts: untouched
    s, a, c = O.barostat_adapt(1.1, 10, 2, 27.0)  # < 25 % accepted: shrink, reset
    assert (a, c) == (0, 0) and s == pytest.approx(1.0)
    s, a, c = O.barostat_adapt(1.0, 10, 8, 27.0)  # > 75 % accepted: grow, reset
    assert (a, c) == (0, 0) and s == pytest.approx(1.1)
    s, _, _ = O.barostat_adapt(8.0, 12, 12, 27.0)  # growth is capped at 30 % of the volume
    assert s == pytest.approx(8.1)
    assert O.barostat_adapt(1.0, 10, 5, 27.0) == (1.0, 10, 5)  # in between: keep counting


def test_get_group_indices(golden):
    groups = O.get_group_indices([tuple(b) for b in golden["bond_list"]], int(golden["num_atoms"]))
    expect = [g for g in golden["groups"] if len(g) > 1] + [g for g in golden["groups"] if len(g) == 1]
    assert len(groups) == len(expect)
    for a, b in zip(groups, expect):
        np.testing.assert_array_equal(a, b)
    # a shuffled edge list gives the same partition
    rng = np.random.default_rng(0)
    bl = [tuple(b[::-1]) for b in golden["bond_list"][rng.permutation(len(golden["bond_list"]))]]
    shuffled = O.get_group_indices(bl, int(golden["num_atoms"]))
    assert sorted(tuple(g) for g in shuffled) == sorted(tuple(g) for g in expect)
