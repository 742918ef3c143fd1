"""CPU: the centroid-restraint and velocity-Verlet oracles (oracle/tm_oracle.py) against golden values computed by the
reference's own Python code (tests/golden/make_golden_verlet_centroid.py -> verlet_centroid.npz), and rmsd_align - a host
function of the library, as in the reference (rmsd_align.cpp) - against a NumPy Kabsch alignment."""

from pathlib import Path

import numpy as np
import pytest

from oracle import tm_oracle as O

G = dict(np.load(Path(__file__).parent / "golden" / "verlet_centroid.npz"))


@pytest.mark.parametrize("tag,b0", [("b0", None), ("zero", 0.0)])
def test_centroid_restraint_oracle(tag, b0):
    b0 = float(G["b0"]) if b0 is None else b0
    u, du_dx = O.centroid_restraint(G["x"], G["group_a"], G["group_b"], float(G["kb"]), b0)
    np.testing.assert_allclose(u, G[f"u_{tag}"], rtol=1e-12)
    np.testing.assert_allclose(du_dx, G[f"du_dx_fd_{tag}"], rtol=1e-6, atol=1e-6)
    # equal and opposite net force on the two groups, none elsewhere
    np.testing.assert_allclose(du_dx[G["group_a"]].sum(0), -du_dx[G["group_b"]].sum(0), rtol=1e-12)
    others = np.setdiff1d(np.arange(len(G["x"])), np.concatenate([G["group_a"], G["group_b"]]))
    assert not np.any(du_dx[others])


def _chain_force(x):
    k2, k4, r0 = float(G["vv_k2"]), float(G["vv_k4"]), float(G["vv_r0"])
    d = x[1:] - x[:-1]
    r = np.linalg.norm(d, axis=1)
    g = ((k2 * (r - r0) + k4 * (r - r0) ** 3) / r)[:, None] * d
    f = np.zeros_like(x)
    f[1:] -= g
    f[:-1] += g
    return f


def test_velocity_verlet_oracle_is_the_references_python_integrator():
    xs, vs = O.velocity_verlet_multiple_steps(
        _chain_force, G["vv_x0"], G["vv_v0"], G["vv_masses"], float(G["vv_dt"]), int(G["vv_n_steps"])
    )
    # both carry x and v in 2^-36 fixed point: identical integers, identical floats
    np.testing.assert_array_equal(xs, G["vv_xs"])
    np.testing.assert_array_equal(vs, G["vv_vs"])


def test_compiled_velocity_verlet_form_tracks_the_python_one():
    """verlet_integrator.cu keeps x and v in f64 without the fixed-point rounding of the Python class: same scheme, so
    after T steps the two agree to the rounding they differ by (T x 2^-36).  The compiled Context runs n steps between its
    two half steps, the Python class n - 1 (tests/test_velocity_verlet_integrator.py:134-135)."""
    n = int(G["vv_n_steps"])
    dt = float(G["vv_dt"])
    cbs = -dt / G["vv_masses"]
    frames, x_end, v_end = O.velocity_verlet_f64(lambda x: -_chain_force(x), G["vv_x0"], G["vv_v0"], cbs, dt, n - 1)
    np.testing.assert_allclose(frames, G["vv_xs"][1:-1], atol=1e-8)
    np.testing.assert_allclose(x_end, G["vv_xs"][-1], atol=1e-8)
    np.testing.assert_allclose(v_end, G["vv_vs"][-1], atol=1e-6)


def _kabsch(x1, x2):
    c1, c2 = x1.mean(0), x2.mean(0)
    a = x2 - c2
    u, _, vt = np.linalg.svd(a.T @ (x1 - c1))
    if np.linalg.det(u) * np.linalg.det(vt) < 0:
        u[:, 2] *= -1
    return a @ (u @ vt) + c1


def test_rmsd_align_against_numpy_kabsch():
    from timemachine_b200 import custom_ops

    rng = np.random.default_rng(5)
    for trial in range(60):
        n = int(rng.integers(3, 50))
        x1 = rng.normal(size=(n, 3)) * rng.uniform(0.1, 5)
        if trial % 3 == 0:  # a rotated, shifted, slightly perturbed copy
            q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            q *= np.sign(np.linalg.det(q))
            x2 = x1 @ q + rng.normal(size=3) + rng.normal(size=(n, 3)) * 0.01
        elif trial % 3 == 1:  # mirror image: the best PROPER rotation is wanted
            x2 = -x1 + rng.normal(size=(n, 3)) * 0.05
        else:
            x2 = rng.normal(size=(n, 3))
        got = custom_ops.rmsd_align(x1, x2)
        np.testing.assert_allclose(got, _kabsch(x1, x2), atol=1e-11)
        np.testing.assert_allclose(got.mean(0), x1.mean(0), atol=1e-12)
    # aligned copy of a rigidly moved structure is the structure
    x1 = rng.normal(size=(20, 3))
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    q *= np.sign(np.linalg.det(q))
    np.testing.assert_allclose(custom_ops.rmsd_align(x1, x1 @ q + 3.0), x1, atol=1e-12)
    # planar and collinear inputs (rank-deficient correlation matrix) stay finite and optimal
    p1, p2 = rng.normal(size=(10, 3)), rng.normal(size=(10, 3))
    p1[:, 2] = p2[:, 2] = 0
    rmsd = lambda a, b: np.sqrt(((a - b) ** 2).sum(1).mean())  # noqa: E731
    assert rmsd(custom_ops.rmsd_align(p1, p2), p1) == pytest.approx(rmsd(_kabsch(p1, p2), p1), abs=1e-12)
    l1, l2 = np.outer(np.arange(5.0), [1, 0, 0]), np.outer(np.arange(5.0), [0, 2, 0])
    assert rmsd(custom_ops.rmsd_align(l1, l2), l1) == pytest.approx(np.sqrt(2.0), abs=1e-12)
    with pytest.raises(RuntimeError, match="N1 != N2"):
        custom_ops.rmsd_align(np.zeros((3, 3)), np.zeros((4, 3)))
    with pytest.raises(RuntimeError, match="D1 != 3"):
        custom_ops.rmsd_align(np.zeros((3, 2)), np.zeros((3, 3)))
