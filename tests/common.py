"""Shared helpers for the parity tests: synthetic systems, comparison conventions, optional access to
the reference's own custom_ops compiled into oracle/_ref (test oracle only)."""

from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
ONE_4PI_EPS0 = 138.935456


def load_reference_ops():
    """The UNMODIFIED reference custom_ops built for sm_100a by oracle/ref_build (None when it has not been built)."""
    d = ROOT / "oracle" / "_ref"
    cands = sorted(d.glob("custom_ops*.so"))
    if not cands:
        import os

        if os.environ.get("TMB_REQUIRE_REF") == "1":
            # the GPU-box runs set this: oracle/_ref travels with the snapshot, and a missing reference must not silently
            # turn the bitwise parity tests into skips
            raise RuntimeError("TMB_REQUIRE_REF=1 but oracle/_ref/custom_ops*.so is not built (make -C oracle/ref_build)")
        return None
    name = "tm_reference_custom_ops"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location("custom_ops", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def require_reference_ops():
    """The compiled reference, or a skip - a hard failure when TMB_REQUIRE_REF=1 (set for the GPU-box runs, where
    oracle/_ref travels with the snapshot: a missing reference must not silently drop the bitwise parity tests)."""
    import os

    import pytest

    ref = load_reference_ops()
    if ref is None:
        if os.environ.get("TMB_REQUIRE_REF") == "1":
            pytest.fail("TMB_REQUIRE_REF=1 but oracle/_ref/custom_ops*.so is not built (make -C oracle/ref_build)")
        pytest.skip("oracle/_ref/custom_ops*.so not built")
    return ref


def assert_forces_close(ref, test, rtol, what="forces"):
    """The reference's convention (tests/common.py:250-273): per atom |dF| <= rtol * max(|F_ref|, 1)."""
    ref = np.asarray(ref)
    test = np.asarray(test)
    assert ref.shape == test.shape
    norms = np.linalg.norm(ref.reshape(len(ref), -1), axis=-1)
    norms = np.where(norms < 1.0, 1.0, norms)
    err = np.linalg.norm((ref - test).reshape(len(ref), -1), axis=-1) / norms
    worst = int(np.argmax(err))
    assert err[worst] <= rtol, f"{what}: atom {worst} relative error {err[worst]:.3e} > {rtol:.1e} (ref {ref[worst]}, got {test[worst]})"


def water_box(n_waters: int, seed: int = 2022, density: float = 33.4, jitter: float = 0.02):
    """TIP3P-like water on a jittered cubic lattice with random orientations, atom order O,H,H.

    Returns dict(x, box, params[N,4] in timemachine encoding (q*sqrt(k_e), sigma/2, sqrt(eps), w), bond/angle idxs and
    params, exclusion pairs/scales, masses (HMR-style)).  Density in molecules / nm^3.
    """
    rng = np.random.default_rng(seed)
    L = (n_waters / density) ** (1.0 / 3.0)
    n_side = int(np.ceil(n_waters ** (1.0 / 3.0)))
    grid = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    pick = rng.permutation(len(grid))[:n_waters]
    centers = (grid[pick] + 0.5) * (L / n_side) + rng.normal(0, jitter, (n_waters, 3))
    # rigid TIP3P geometry, random rotation per molecule
    r_oh, theta = 0.09572, 1.82421813
    h1 = np.array([r_oh, 0.0, 0.0])
    h2 = np.array([r_oh * np.cos(theta), r_oh * np.sin(theta), 0.0])
    qn = rng.normal(size=(n_waters, 4))
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    a, b, c, d = qn.T
    R = np.stack(
        [
            np.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)], -1),
            np.stack([2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)], -1),
            np.stack([2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1),
        ],
        axis=1,
    )
    x = np.empty((n_waters, 3, 3))
    x[:, 0] = centers
    x[:, 1] = centers + R @ h1
    x[:, 2] = centers + R @ h2
    x = x.reshape(-1, 3)
    N = 3 * n_waters
    sq = np.sqrt(ONE_4PI_EPS0)
    params = np.zeros((N, 4))
    params[0::3] = [-0.834 * sq, 0.1575375, np.sqrt(0.635968), 0.0]
    params[1::3] = [0.417 * sq, 0.05, 0.0, 0.0]
    params[2::3] = [0.417 * sq, 0.05, 0.0, 0.0]
    o = np.arange(0, N, 3, dtype=np.int32)
    bond_idxs = np.concatenate([np.stack([o, o + 1], 1), np.stack([o, o + 2], 1)]).astype(np.int32)
    bond_params = np.tile([462750.4, r_oh], (len(bond_idxs), 1))
    angle_idxs = np.stack([o + 1, o, o + 2], 1).astype(np.int32)
    angle_params = np.tile([836.8, theta, 0.0], (len(angle_idxs), 1))
    excl = np.concatenate([np.stack([o, o + 1], 1), np.stack([o, o + 2], 1), np.stack([o + 1, o + 2], 1)]).astype(np.int32)
    scales = np.ones((len(excl), 2))
    masses = np.tile([15.999 - 2 * 2.016, 1.008 + 2.016, 1.008 + 2.016], n_waters)
    return dict(
        x=x, box=np.eye(3) * L, params=params, bond_idxs=bond_idxs, bond_params=bond_params, angle_idxs=angle_idxs,
        angle_params=angle_params, exclusion_idxs=excl, scale_factors=scales, masses=masses, N=N,
    )


def random_nonbonded_system(n: int, seed: int, box_len: float | None = None, w_pattern: str = "zero", density: float = 100.0):
    """Generic charged LJ particles at liquid-like density without overlaps (jittered lattice) for kernel parity tests.
    w_pattern as the reference's gen_nonbonded_params_with_4d_offsets (tests/common.py:348-379)."""
    rng = np.random.default_rng(seed)
    if box_len is None:
        box_len = max((n / density) ** (1 / 3), 2.7)
    n_side = int(np.ceil(n ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    pick = rng.permutation(len(grid))[:n]
    x = (grid[pick] + 0.5) * (box_len / n_side) + rng.normal(0, 0.02, (n, 3))
    # scatter some atoms into other periodic images: results must not care
    shift = rng.integers(-2, 3, (n, 3)) * (rng.random((n, 1)) < 0.25)
    x = x + shift * box_len
    q = rng.normal(0, 0.4, n) * np.sqrt(ONE_4PI_EPS0)
    sig = rng.uniform(0.1, 0.17, n)
    eps = np.sqrt(rng.uniform(0.05, 1.0, n))
    eps[rng.random(n) < 0.3] = 0.0
    w = np.zeros(n)
    cutoff = 1.2
    if w_pattern == "zero":
        pass
    elif w_pattern == "all_same":
        w[:] = 0.3
    elif w_pattern == "some":
        m = rng.random(n) < 0.2
        w[m] = rng.uniform(-0.5 * cutoff, 0.5 * cutoff, m.sum())
    elif w_pattern == "cutoff":
        m = rng.random(n) < 0.2
        w[m] = cutoff  # parked exactly at the cutoff: must not interact with w == 0 atoms
    elif w_pattern == "beyond":
        m = rng.random(n) < 0.2
        w[m] = rng.uniform(cutoff, 2 * cutoff, m.sum())
    else:
        raise ValueError(w_pattern)
    params = np.stack([q, sig, eps, w], 1)
    return x, params, np.eye(3) * box_len


def round_to_f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)
