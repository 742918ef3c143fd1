"""GPU: BASELINE.json's configurations as tests (the bench line itself is configs[3]'s per-GPU leg; configs[1] is
tests/test_scale_gpu.py's DHFR-sized box).

  configs[0]  256-atom vacuum box, HarmonicBond + Nonbonded, 100 Langevin steps against the NumPy oracle (the CPU
              restatement of the reference's JAX potentials + its BAOAB integrator)
  configs[2]  solvated ligand, 4D-interpolated interaction group at lambda = 0.5: du/dp and du/dl
  configs[4]  90k-atom stress box: device-side rebuilds keep firing under MD, graph replay == eager stepping, exact
              fixed-point properties at full size
"""

import numpy as np
import pytest

from oracle import tm_oracle as O
from tests.common import assert_forces_close, round_to_f32, water_box

pytestmark = pytest.mark.gpu
BETA, CUTOFF = 2.0, 1.2


def mods():
    from timemachine_b200 import custom_ops, lib, potentials

    return custom_ops, lib, potentials


# ---------------------------------------------------------------------------------------------------------------------
def _vacuum_cluster():
    s = water_box(85, seed=4)
    # 85 waters + one ion = 256 atoms, in a box far larger than the cluster: no image ever interacts (tests/test_bonded.py:98)
    x = np.concatenate([s["x"], s["x"].mean(axis=0, keepdims=True) + [[0.9, 0.0, 0.0]]])
    params = np.concatenate([s["params"], [[1.0 * np.sqrt(138.935456), 0.12, np.sqrt(0.4), 0.0]]])
    masses = np.concatenate([s["masses"], [22.99]])
    s.update(x=round_to_f32(x), params=round_to_f32(params), masses=masses, N=256, box=np.eye(3) * 100.0)
    return s


@pytest.mark.parametrize("precision,atol", [(np.float64, 2e-7), (np.float32, 2e-4)])
def test_config0_vacuum_256_atoms_100_langevin_steps(precision, atol):
    ops, lib, P = mods()
    s = _vacuum_cluster()
    N = s["N"]
    assert N == 256
    rng = np.random.default_rng(0)
    v0 = rng.normal(0, 1.0, (N, 3)) * np.sqrt(0.008314462618 * 300.0 / s["masses"])[:, None]
    dt, temperature, friction = 1.0e-3, 300.0, 0.0  # friction 0: the noise coefficient vanishes (tests/test_md.py:174-178)
    bond = P.HarmonicBond(s["bond_idxs"]).bind(s["bond_params"]).to_gpu(precision).bound_impl
    nb = P.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF).bind(s["params"]).to_gpu(precision).bound_impl
    intg = lib.LangevinIntegrator(temperature, dt, friction, s["masses"], 2024).impl()
    ctx = ops.Context(s["x"], v0, s["box"], intg, [bond, nb])
    xs, boxes = ctx.multiple_steps(100, 25)
    assert xs.shape == (4, N, 3)

    ca, cb, cc = O.langevin_coefficients(temperature, dt, friction, s["masses"])
    x, v = s["x"].copy(), v0.copy()
    frames = []
    for step in range(1, 101):
        f = -(
            O.harmonic_bond(x, s["bond_params"], s["bond_idxs"])[1]
            + O.nonbonded(x, s["params"], s["box"], s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF)[1]
        )
        x, v = O.baoab_step(x, v, f, ca, cb, cc, dt, np.zeros_like(x))
        if step % 25 == 0:
            frames.append(x.copy())
    # the integrator arithmetic is f32 on f64 state in both precisions (reference k_integrator.cuh:32-46); the f64
    # potentials leave only that, the f32 potentials add their 1e-4-relative force error over 100 steps
    np.testing.assert_allclose(xs, np.array(frames), rtol=0, atol=max(atol, 5e-6))
    assert np.abs(xs[-1] - s["x"]).max() > 0.02  # the cluster really moved


# ---------------------------------------------------------------------------------------------------------------------
def _solvated_ligand(n_waters=700, n_lig=24, seed=9):
    rng = np.random.default_rng(seed)
    s = water_box(n_waters, seed=seed)
    L = s["box"][0, 0]
    lig = L / 2 + rng.normal(0, 0.25, (n_lig, 3))
    n_env = s["N"]
    N = n_env + n_lig
    x = np.concatenate([s["x"], lig])
    params = np.zeros((N, 4))
    params[:n_env] = s["params"]
    params[n_env:, 0] = rng.normal(0, 0.3, n_lig) * np.sqrt(138.935456)
    params[n_env:, 1] = rng.uniform(0.12, 0.18, n_lig)
    params[n_env:, 2] = np.sqrt(rng.uniform(0.2, 0.6, n_lig))
    return dict(x=round_to_f32(x), box=s["box"], params=params, N=N, n_env=n_env, lig=np.arange(n_env, N, dtype=np.int32))


def _params_at(s, lam):
    """Half of the ligand is decoupled through the 4th dimension and its charges are scaled: w = lam * cutoff,
    q = (1 - lam) q0 (the shape of fe/single_topology.py:934-951's interpolation)."""
    p = s["params"].copy()
    dummy = s["lig"][len(s["lig"]) // 2 :]
    p[dummy, 3] = lam * CUTOFF
    p[dummy, 0] *= 1.0 - lam
    return p


def test_config2_interaction_group_du_dp_and_du_dl_at_lambda_half():
    ops, lib, P = mods()
    s = _solvated_ligand()
    N, lam, h = s["N"], 0.5, 1e-5
    ixn = P.NonbondedInteractionGroup(N, s["lig"], BETA, CUTOFF)
    g64 = ixn.to_gpu(np.float64).unbound_impl
    g32 = ixn.to_gpu(np.float32).unbound_impl
    p = _params_at(s, lam)
    dx64, dp64, u64 = g64.execute(s["x"], p, s["box"])
    dx32, dp32, u32 = g32.execute(s["x"], round_to_f32(p), s["box"])
    # against the oracle (rows = ligand, cols = environment)
    env = np.arange(s["n_env"])
    ref_u, ref_dx, ref_dp = O.nonbonded_interaction_group(s["x"], p, s["box"], s["lig"], env, BETA, CUTOFF)
    np.testing.assert_allclose(u64, ref_u, rtol=1e-9)
    assert_forces_close(ref_dx, dx64, 1e-8)
    assert_forces_close(ref_dp, dp64, 1e-7, what="du_dp")
    np.testing.assert_allclose(u32, ref_u, rtol=2e-4, atol=5e-3)
    assert_forces_close(ref_dx, dx32, 1e-4)
    assert_forces_close(ref_dp, dp32, 1e-3, what="du_dp")
    assert np.any(dp64[:, 3] != 0)  # the w column carries the coupling
    # du/dl = sum_p du/dp * dp/dl (the reference chains through JAX, there is no du_dl in the C++ API)
    dp_dl = (_params_at(s, lam + h) - _params_at(s, lam - h)) / (2 * h)
    du_dl = float(np.sum(dp64 * dp_dl))
    u_plus = g64.execute(s["x"], _params_at(s, lam + h), s["box"], False, False, True)[2]
    u_minus = g64.execute(s["x"], _params_at(s, lam - h), s["box"], False, False, True)[2]
    fd = (u_plus - u_minus) / (2 * h)
    np.testing.assert_allclose(du_dl, fd, rtol=2e-5, atol=1e-4)
    np.testing.assert_allclose(float(np.sum(dp32 * dp_dl)), fd, rtol=5e-3, atol=5e-2)  # f32 kernels, same chain rule


# ---------------------------------------------------------------------------------------------------------------------
def test_config4_stress_box_rebuilds_under_md_and_replays_bitwise():
    ops, lib, P = mods()
    s = water_box(30000, seed=31)  # 90,000 atoms, box 9.65 nm
    N = s["N"]
    assert N == 90000
    x = round_to_f32(s["x"])
    rng = np.random.default_rng(1)
    v0 = rng.normal(0, 1.0, (N, 3)) * np.sqrt(0.008314462618 * 300.0 / s["masses"])[:, None]
    flat = np.concatenate([s["bond_params"].reshape(-1), s["angle_params"].reshape(-1), s["params"].reshape(-1)])

    def run(use_graphs):
        nb = P.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF)
        pot = P.SummedPotential(
            [P.HarmonicBond(s["bond_idxs"]), P.HarmonicAngle(s["angle_idxs"]), nb], [s["bond_params"], s["angle_params"], s["params"]]
        )
        impl = pot.to_gpu(np.float32).unbound_impl
        all_pairs = impl.get_potentials()[2].get_potentials()[0]
        intg = lib.LangevinIntegrator(300.0, 1.0e-3, 1.0, s["masses"], 77).impl()
        ctx = ops.Context(x, v0, s["box"], intg, [ops.BoundPotential(impl, flat)])
        ctx.set_use_graphs(use_graphs)
        xs, _ = ctx.multiple_steps(64)
        return xs[-1], ctx.get_v_t(), all_pairs.get_num_rebuilds(), all_pairs.get_tile_count(), impl

    xa, va, rebuilds_a, tiles, impl = run(True)
    xb, vb, rebuilds_b, _, _ = run(False)
    assert np.isfinite(xa).all()
    # the rebuild decision lives on the device: it fires the same way whether the steps are replayed from a CUDA graph
    # or launched one by one, and the trajectories are bitwise equal (Philox noise is keyed by step, not by launch)
    assert rebuilds_a == rebuilds_b and rebuilds_a >= 64 // 16
    np.testing.assert_array_equal(xa, xb)
    np.testing.assert_array_equal(va, vb)
    assert 0.8 * N < tiles < 1.6 * N  # about one 32 x 32 tile per atom at this density
    # exact property at full size on the final frame: Newton's third law of the pair terms holds in fixed point (every
    # pair adds +v to one atom and exactly -v to the other; the angle term rounds its three forces separately)
    nonbonded = impl.get_potentials()[2]
    du_dx, _, u = nonbonded.execute(xa, s["params"], s["box"], True, False, True)
    assert np.isfinite(u)
    assert not np.rint(du_dx * 2.0**36).astype(np.int64).sum(axis=0).any()
