"""CPU: the C-ABI library loads and exports every symbol include/tmb200.h declares; the ctypes table covers them all.
No compute call is made here (there is no GPU in the build container)."""

import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "tmb200.h"


def declared_symbols():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tmb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_a_sane_number_of_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 60
    for must in ("tmb_potential_execute", "tmb_context_multiple_steps", "tmb_nonbonded_all_pairs_create", "tmb_hilbert_sort_sort"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from timemachine_b200 import _lib

    assert _lib.LIB_PATH.exists(), "build the library first: python -m timemachine_b200.build"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in tmb200.h but not exported: {missing}"


def test_ctypes_table_matches_header():
    from timemachine_b200 import _lib

    declared = set(declared_symbols())
    table = set(_lib.SIGNATURES)
    assert declared == table, f"header-only: {sorted(declared - table)}; table-only: {sorted(table - declared)}"


def test_custom_ops_surface_matches_reference_names():
    """The shim must expose the reference's hot-path class names (timemachine/lib/custom_ops.pyi)."""
    from timemachine_b200 import custom_ops as ops

    names = [
        "Potential", "BoundPotential", "SummedPotential", "FanoutSummedPotential", "Integrator", "LangevinIntegrator",
        "Context", "HilbertSort", "InvalidHardware", "FIXED_EXPONENT", "cuda_device_reset",
    ]
    for base in ("HarmonicBond", "HarmonicAngle", "PeriodicTorsion", "NonbondedAllPairs", "NonbondedInteractionGroup",
                 "NonbondedPairList", "NonbondedExclusions", "Neighborlist"):
        names += [f"{base}_f32", f"{base}_f64"]
    for n in names:
        assert hasattr(ops, n), n
    assert ops.FIXED_EXPONENT == 2**36
    for cls in (ops.HarmonicBond_f32, ops.NonbondedAllPairs_f64, ops.SummedPotential):
        assert issubclass(cls, ops.Potential)
    for m in ("execute", "execute_du_dx", "execute_batch", "execute_batch_sparse"):
        assert hasattr(ops.Potential, m)
    for m in ("step", "multiple_steps", "set_x_t", "get_x_t", "set_v_t", "get_v_t", "set_box", "get_box", "initialize", "finalize"):
        assert hasattr(ops.Context, m)


def test_host_side_validation_without_gpu():
    """Argument checks that run before any CUDA call keep the reference's messages (wrap_kernels.cpp:51-78)."""
    import numpy as np

    from timemachine_b200 import custom_ops as ops

    with pytest.raises(RuntimeError, match="coords dimensions must be 2"):
        ops._verify_coords_and_box(np.zeros((2, 3, 3)), np.eye(3))
    with pytest.raises(RuntimeError, match="box must be 3x3"):
        ops._verify_coords_and_box(np.zeros((2, 3)), np.eye(2))
    with pytest.raises(RuntimeError, match="box must have positive values along diagonal"):
        ops._verify_coords_and_box(np.zeros((2, 3)), np.zeros((3, 3)))
    bad = np.eye(3)
    bad[0, 1] = 0.1
    with pytest.raises(RuntimeError, match="box must be ortholinear"):
        ops._verify_coords_and_box(np.zeros((2, 3)), bad)


def test_exchange_surface_and_host_side_validation_without_gpu():
    """the exchange movers and their helpers are exported under the reference's names
    (timemachine/lib/custom_ops.pyi), and the argument checks that precede any CUDA call keep its messages
    (wrap_kernels.cpp:1749-1755, 2055-2100, 2124-2126)."""
    import numpy as np

    from timemachine_b200 import custom_ops as ops

    for base in ("BDExchangeMove", "TIBDExchangeMove", "NonbondedMolEnergyPotential", "SegmentedSumExp", "SegmentedWeightedRandomSampler",
                 "atom_by_atom_energies", "rotate_coords", "rotate_and_translate_mol", "inner_and_outer_mols",
                 "translations_inside_and_outside_sphere_host"):
        for suffix in ("f32", "f64"):
            assert hasattr(ops, f"{base}_{suffix}"), f"{base}_{suffix}"
    assert issubclass(ops.BDExchangeMove_f32, ops.Mover) and issubclass(ops.TIBDExchangeMove_f64, ops.BDExchangeMove_f64)
    for m in ("move", "compute_initial_log_weights", "compute_incremental_log_weights", "get_params", "set_params", "last_log_probability",
              "last_raw_log_probability", "n_accepted", "n_proposed", "acceptance_fraction", "get_before_log_weights",
              "get_after_log_weights", "batch_size", "set_interval", "get_interval", "set_step"):
        assert hasattr(ops.BDExchangeMove_f32, m), m
    params = np.zeros((10, 4))
    mols = [[0, 1, 2], [3, 4, 5]]
    for k, extra in ((ops.BDExchangeMove_f32, ()), (ops.TIBDExchangeMove_f32, ([9],))):
        lead = (10, *extra, mols)
        tail = (300.0, 2.0, 1.2) + ((1.0,) if extra else ())
        with pytest.raises(RuntimeError, match="proposals per move must be greater than 0"):
            k(*lead, params, *tail, 1, 0, 1)
        with pytest.raises(RuntimeError, match="parameters dimensions must be 2"):
            k(*lead, params.reshape(-1), *tail, 1, 1, 1)
        with pytest.raises(RuntimeError, match="Number of parameters must match N"):
            k(*lead, params[:9], *tail, 1, 1, 1)
    with pytest.raises(RuntimeError, match="quaternions must have a shape that is 4 dimensional"):
        ops.rotate_coords_f32(np.zeros((3, 3)), np.zeros((2, 3)))
    with pytest.raises(RuntimeError, match="quaternions dimensions must be 2"):
        ops.rotate_and_translate_mol_f64(np.zeros((3, 3)), np.eye(3), np.zeros(4), np.zeros((1, 3)))
    with pytest.raises(RuntimeError, match="translations must be of size 3"):
        ops.rotate_and_translate_mol_f64(np.zeros((3, 3)), np.eye(3), np.zeros((1, 4)), np.zeros((1, 2)))
    with pytest.raises(RuntimeError, match="Number of quaternions and translations must match"):
        ops.rotate_and_translate_mol_f64(np.zeros((3, 3)), np.eye(3), np.zeros((2, 4)), np.zeros((1, 3)))
    with pytest.raises(RuntimeError, match="Center must be of length 3"):
        ops.translations_inside_and_outside_sphere_host_f32(4, np.eye(3), np.zeros(2), 1.0, 1)
