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
