"""Drop-in for the hot-path subset of `timemachine.lib.custom_ops` (reference: timemachine/cpp/src/wrap_kernels.cpp,
API listing timemachine/lib/custom_ops.pyi), backed by the sm_100a CUDA library libtmb200.so through its C ABI.

Same class names, constructor argument order, method names, return shapes, fixed-point -> float conversion and error
strings as the reference module, so the reference's own Python wrappers
(`timemachine.potentials.Potential.to_gpu`, `timemachine.lib.LangevinIntegrator.impl`, `custom_ops.Context`) can run on
top of it unchanged:

    import sys, timemachine_b200.custom_ops as ops
    sys.modules["timemachine.lib.custom_ops"] = ops        # before importing timemachine.potentials

There is no CPU fallback: importing this module without the built CUDA library raises ImportError.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import F32, F64, I128

_L = _lib.load()

FIXED_EXPONENT = int(_L.tmb_fixed_exponent())  # wrap_kernels.cpp:2311
_LLONG_MAX = (1 << 63) - 1
_LLONG_MIN = -(1 << 63)


class InvalidHardware(Exception):
    """Raised when the CUDA device/driver is unusable (reference exceptions.hpp, wrap_kernels.cpp:2144)."""


def _check(status: int) -> None:
    if status == _lib.TMB_OK:
        return
    msg = (_L.tmb_last_error() or b"").decode("utf-8", "replace")
    if status == _lib.TMB_INVALID_HARDWARE:
        raise InvalidHardware(msg)
    raise RuntimeError(msg)


def cuda_device_reset() -> None:
    _check(_L.tmb_cuda_device_reset())


def kernel_launch_count() -> int:
    """Kernels launched by libtmb200 so far (bench.py reports the delta over the timed region)."""
    return int(_L.tmb_kernel_launch_count())


# ---- marshalling helpers -------------------------------------------------------------------------------------------
def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


def _verify_coords(coords: np.ndarray) -> None:
    if coords.ndim != 2:
        raise RuntimeError("coords dimensions must be 2")
    if coords.shape[-1] != 3:
        raise RuntimeError("coords must have a shape that is 3 dimensional")


def _verify_coords_and_box(coords: np.ndarray, box: np.ndarray) -> None:
    """wrap_kernels.cpp:51-78"""
    _verify_coords(coords)
    if box.ndim != 2 or box.shape[0] != 3 or box.shape[1] != 3:
        raise RuntimeError("box must be 3x3")
    flat = box.reshape(-1)
    for i in range(9):
        if i in (0, 4, 8):
            if flat[i] <= 0.0:
                raise RuntimeError("box must have positive values along diagonal")
        elif flat[i] != 0.0:
            raise RuntimeError("box must be ortholinear")


def _fixed_to_float(fixed: np.ndarray) -> np.ndarray:
    """FIXED_TO_FLOAT<double>: (double)(int64)v / 2^36 (fixed_point.hpp:17-19)"""
    return fixed.view(np.int64).astype(np.float64) / float(FIXED_EXPONENT)


def _energy_to_float(u: I128) -> float:
    """convert_energy_to_fp (wrap_kernels.cpp:83-89): NaN when the int128 sum left the int64 range."""
    v = (int(u.hi) << 64) | int(u.lo)
    if v >= _LLONG_MAX or v <= _LLONG_MIN:
        return float("nan")
    return float(np.float64(v) / np.float64(FIXED_EXPONENT))


def _energies_to_float(us) -> np.ndarray:
    return np.array([_energy_to_float(u) for u in us], dtype=np.float64)


# ---- Potential -----------------------------------------------------------------------------------------------------
class Potential:
    """Base class of every potential (wrap_kernels.cpp:729-1131). Holds a tmb_potential handle."""

    _handle: Optional[C.c_void_p] = None

    def __init__(self, *args, **kwargs):
        raise TypeError("Potential has no constructor; use one of the concrete potentials")

    def _adopt(self, handle: C.c_void_p) -> None:
        self._handle = handle

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and _L is not None:
            try:
                _L.tmb_potential_destroy(h)
            except Exception:
                pass
            self._handle = None

    # -- single evaluation ------------------------------------------------------------------------------------------
    def execute(self, coords, params, box, compute_du_dx=True, compute_du_dp=True, compute_u=True):
        coords = _f64(coords)
        params = _f64(params)
        box = _f64(box)
        _verify_coords_and_box(coords, box)
        N = coords.shape[0]
        P = params.size
        du_dx = np.empty((N, 3), dtype=np.uint64) if compute_du_dx else None
        du_dp = np.empty(P, dtype=np.uint64) if compute_du_dp else None
        u = I128() if compute_u else None
        _check(
            _L.tmb_potential_execute(
                self._handle, N, P, _ptr(coords, C.c_double), _ptr(params, C.c_double), _ptr(box, C.c_double),
                _ptr(du_dx, C.c_uint64), _ptr(du_dp, C.c_uint64), C.byref(u) if compute_u else None,
            )
        )
        out_dx = _fixed_to_float(du_dx) if compute_du_dx else None
        out_dp = None
        if compute_du_dp:
            out_dp = np.empty(params.shape, dtype=np.float64)
            if P > 0:
                _check(_L.tmb_potential_du_dp_fixed_to_float(self._handle, N, P, _ptr(du_dp, C.c_uint64), _ptr(out_dp, C.c_double)))
        out_u = _energy_to_float(u) if compute_u else None
        return out_dx, out_dp, out_u

    def execute_du_dx(self, coords, params, box):
        return self.execute(coords, params, box, True, False, False)[0]

    # -- dense batch: every coords x every params --------------------------------------------------------------------
    def execute_batch(self, coords, params, boxes, compute_du_dx, compute_du_dp, compute_u):
        coords = _f64(coords)
        params = _f64(params)
        boxes = _f64(boxes)
        if coords.ndim != 3 or boxes.ndim != 3:
            raise RuntimeError("coords and boxes must have 3 dimensions")
        if coords.shape[0] != boxes.shape[0]:
            raise RuntimeError("number of batches of coords and boxes don't match")
        if params.ndim < 2:
            raise RuntimeError("parameters must have at least 2 dimensions")
        CB, N, D = coords.shape
        PB = params.shape[0]
        P = params.size // PB if PB else 0
        total = CB * PB
        du_dx = np.empty((total, N, D), dtype=np.uint64) if compute_du_dx else None
        du_dp = np.empty((total, P), dtype=np.uint64) if compute_du_dp else None
        u = (I128 * total)() if compute_u else None
        _check(
            _L.tmb_potential_execute_batch(
                self._handle, CB, N, PB, P, _ptr(coords, C.c_double), _ptr(params, C.c_double), _ptr(boxes, C.c_double),
                _ptr(du_dx, C.c_uint64), _ptr(du_dp, C.c_uint64), u,
            )
        )
        out_dx = _fixed_to_float(du_dx).reshape(CB, PB, N, D) if compute_du_dx else None
        out_dp = None
        if compute_du_dp:
            out_dp = np.empty((CB, PB) + tuple(params.shape[1:]), dtype=np.float64)
            flat = out_dp.reshape(total, P)
            for i in range(total):
                if P > 0:
                    _check(_L.tmb_potential_du_dp_fixed_to_float(
                        self._handle, N, P, _ptr(du_dp[i], C.c_uint64), flat[i].ctypes.data_as(C.POINTER(C.c_double))))
        out_u = _energies_to_float(u).reshape(CB, PB) if compute_u else None
        return out_dx, out_dp, out_u

    # -- sparse batch: explicit (coords_idx, params_idx) pairs ------------------------------------------------------------
    def execute_batch_sparse(
        self, coords, params, boxes, coords_batch_idxs, params_batch_idxs, compute_du_dx, compute_du_dp, compute_u
    ):
        coords = _f64(coords)
        params = _f64(params)
        boxes = _f64(boxes)
        cidx = _u32(coords_batch_idxs)
        pidx = _u32(params_batch_idxs)
        if coords.ndim != 3 or boxes.ndim != 3:
            raise RuntimeError("coords and boxes must have 3 dimensions")
        if coords.shape[0] != boxes.shape[0]:
            raise RuntimeError("number of coord arrays and boxes don't match")
        if params.ndim < 2:
            raise RuntimeError("parameters must have at least 2 dimensions")
        if cidx.ndim != 1 or pidx.ndim != 1:
            raise RuntimeError("coords_batch_idxs and params_batch_idxs must be one-dimensional arrays")
        if cidx.size != pidx.size:
            raise RuntimeError("coords_batch_idxs and params_batch_idxs must have the same length")
        if cidx.size and cidx.max() >= coords.shape[0]:
            raise RuntimeError("coords_batch_idxs contains an index that is out of bounds")
        if pidx.size and pidx.max() >= params.shape[0]:
            raise RuntimeError("params_batch_idxs contains an index that is out of bounds")
        CS, N, D = coords.shape
        PS = params.shape[0]
        P = params.size // PS if PS else 0
        B = cidx.size
        du_dx = np.empty((B, N, D), dtype=np.uint64) if compute_du_dx else None
        du_dp = np.empty((B, P), dtype=np.uint64) if compute_du_dp else None
        u = (I128 * B)() if compute_u else None
        _check(
            _L.tmb_potential_execute_batch_sparse(
                self._handle, CS, N, PS, P, B, _ptr(cidx, C.c_uint32), _ptr(pidx, C.c_uint32), _ptr(coords, C.c_double),
                _ptr(params, C.c_double), _ptr(boxes, C.c_double), _ptr(du_dx, C.c_uint64), _ptr(du_dp, C.c_uint64), u,
            )
        )
        out_dx = _fixed_to_float(du_dx) if compute_du_dx else None
        out_dp = None
        if compute_du_dp:
            out_dp = np.empty((B,) + tuple(params.shape[1:]), dtype=np.float64)
            flat = out_dp.reshape(B, P)
            for i in range(B):
                if P > 0:
                    _check(_L.tmb_potential_du_dp_fixed_to_float(
                        self._handle, N, P, _ptr(du_dp[i], C.c_uint64), flat[i].ctypes.data_as(C.POINTER(C.c_double))))
        out_u = _energies_to_float(u) if compute_u else None
        return out_dx, out_dp, out_u

    # -- device-pointer entry (no reference Python equivalent; Potential::execute_device, potential.hpp:87-96) ---------
    def execute_device(self, N, P, d_coords, d_params, d_box, d_du_dx=0, d_du_dp=0, d_u=0, stream=0) -> None:
        """Raw CUDA pointers (ints). Accumulates into du_dx/du_dp, overwrites u, enqueues on `stream`, no sync."""
        _check(
            _L.tmb_potential_execute_device(
                self._handle, N, P, C.c_void_p(d_coords), C.c_void_p(d_params or None), C.c_void_p(d_box),
                C.c_void_p(d_du_dx or None), C.c_void_p(d_du_dp or None), C.c_void_p(d_u or None), C.c_void_p(stream or None),
            )
        )


def _new_handle() -> C.c_void_p:
    return C.c_void_p()


def _make_bonded(name: str, create_fn, precision: int, arg: str):
    """One-argument constructors; `arg` is the reference's keyword for it (custom_ops.pyi)."""

    def init(self, idxs):
        idxs = _i32(idxs)
        h = _new_handle()
        _check(create_fn(precision, _ptr(idxs, C.c_int32), idxs.size, C.byref(h)))
        self._adopt(h)

    ns = {"init": init}
    exec(f"def __init__(self, {arg}):\n    init(self, {arg})\n", ns)
    return type(name, (Potential,), {"__init__": ns["__init__"], "__doc__": f"{name}({arg}) - see include/tmb200.h"})


HarmonicBond_f32 = _make_bonded("HarmonicBond_f32", _L.tmb_harmonic_bond_create, F32, "bond_idxs")
HarmonicBond_f64 = _make_bonded("HarmonicBond_f64", _L.tmb_harmonic_bond_create, F64, "bond_idxs")
HarmonicAngle_f32 = _make_bonded("HarmonicAngle_f32", _L.tmb_harmonic_angle_create, F32, "angle_idxs")
HarmonicAngle_f64 = _make_bonded("HarmonicAngle_f64", _L.tmb_harmonic_angle_create, F64, "angle_idxs")
# "angle_idxs" is the reference's keyword for the torsion indices too (wrap_kernels.cpp:1439-1444)
PeriodicTorsion_f32 = _make_bonded("PeriodicTorsion_f32", _L.tmb_periodic_torsion_create, F32, "angle_idxs")
PeriodicTorsion_f64 = _make_bonded("PeriodicTorsion_f64", _L.tmb_periodic_torsion_create, F64, "angle_idxs")


FlatBottomBond_f32 = _make_bonded("FlatBottomBond_f32", _L.tmb_flat_bottom_bond_create, F32, "bond_idxs")
FlatBottomBond_f64 = _make_bonded("FlatBottomBond_f64", _L.tmb_flat_bottom_bond_create, F64, "bond_idxs")
ChiralAtomRestraint_f32 = _make_bonded("ChiralAtomRestraint_f32", _L.tmb_chiral_atom_restraint_create, F32, "idxs")
ChiralAtomRestraint_f64 = _make_bonded("ChiralAtomRestraint_f64", _L.tmb_chiral_atom_restraint_create, F64, "idxs")


class _LogFlatBottomBond(Potential):
    """LogFlatBottomBond_{f32,f64}(bond_idxs[B,2], beta): -log(1 - exp(-beta u_flat_bottom)) / beta; params [B,3] = (k, r_min,
    r_max) (wrap_kernels.cpp:1337-1349)."""

    _precision = F32

    def __init__(self, bond_idxs, beta):
        idxs = _i32(bond_idxs)
        h = _new_handle()
        _check(_L.tmb_log_flat_bottom_bond_create(self._precision, _ptr(idxs, C.c_int32), idxs.size, float(beta), C.byref(h)))
        self._adopt(h)


class LogFlatBottomBond_f32(_LogFlatBottomBond):
    _precision = F32


class LogFlatBottomBond_f64(_LogFlatBottomBond):
    _precision = F64


class _CentroidRestraint(Potential):
    """CentroidRestraint_{f32,f64}(group_a_idxs, group_b_idxs, kb, b0): kb (|centroid_a - centroid_b| - b0)^2, geometric
    centroids, no parameters (wrap_kernels.cpp:1410-1430, potentials/bonded.py:8-31)."""

    def __init__(self, group_a_idxs, group_b_idxs, kb, b0):
        ga = _i32(group_a_idxs).reshape(-1)
        gb = _i32(group_b_idxs).reshape(-1)
        h = _new_handle()
        _check(
            _L.tmb_centroid_restraint_create(
                self._precision, _ptr(ga, C.c_int32), ga.size, _ptr(gb, C.c_int32), gb.size, float(kb), float(b0), C.byref(h)
            )
        )
        self._handle = h


class CentroidRestraint_f32(_CentroidRestraint):
    _precision = F32


class CentroidRestraint_f64(_CentroidRestraint):
    _precision = F64


class _ChiralBondRestraint(Potential):
    """ChiralBondRestraint_{f32,f64}(idxs[R,4], signs[R]) (wrap_kernels.cpp:1380-1394)."""

    _precision = F32

    def __init__(self, idxs, signs):
        idxs, signs = _i32(idxs), _i32(signs)
        h = _new_handle()
        _check(
            _L.tmb_chiral_bond_restraint_create(
                self._precision, _ptr(idxs, C.c_int32), idxs.size, _ptr(signs, C.c_int32), signs.size, C.byref(h)
            )
        )
        self._adopt(h)


class ChiralBondRestraint_f32(_ChiralBondRestraint):
    _precision = F32


class ChiralBondRestraint_f64(_ChiralBondRestraint):
    _precision = F64


class _NonbondedPairListPrecomputed(Potential):
    """NonbondedPairListPrecomputed_{f32,f64}(pair_idxs[M,2], beta, cutoff); params are per pair
    (q_ij, sig_ij, eps_ij, w_offset_ij) (wrap_kernels.cpp:1351-1364)."""

    _precision = F32

    def __init__(self, pair_idxs, beta, cutoff):
        pair_idxs = _i32(pair_idxs)
        h = _new_handle()
        _check(
            _L.tmb_nonbonded_pair_list_precomputed_create(
                self._precision, _ptr(pair_idxs, C.c_int32), pair_idxs.size, float(beta), float(cutoff), C.byref(h)
            )
        )
        self._adopt(h)


class NonbondedPairListPrecomputed_f32(_NonbondedPairListPrecomputed):
    _precision = F32


class NonbondedPairListPrecomputed_f64(_NonbondedPairListPrecomputed):
    _precision = F64


class _TiledExtras:
    """Introspection / measurement hooks of the tile-list potentials (not part of the reference API)."""

    def get_tile_count(self) -> int:
        n = C.c_uint()
        _check(_L.tmb_nonbonded_num_tiles(self._handle, C.byref(n)))
        return n.value

    def get_num_rebuilds(self) -> int:
        n = C.c_uint()
        _check(_L.tmb_nonbonded_num_rebuilds(self._handle, C.byref(n)))
        return n.value

    def get_tile_capacity(self):
        """(tiles the list buffer holds now, worst-case tiles the reference would allocate)"""
        cap, worst = C.c_ulonglong(), C.c_ulonglong()
        _check(_L.tmb_nonbonded_tile_capacity(self._handle, C.byref(cap), C.byref(worst)))
        return cap.value, worst.value

    def set_kernel_timing(self, on: bool) -> None:
        _check(_L.tmb_nonbonded_set_kernel_timing(self._handle, int(bool(on))))

    def drain_kernel_times(self) -> np.ndarray:
        out = np.empty(4096, dtype=np.float32)
        n = C.c_int()
        _check(_L.tmb_nonbonded_drain_kernel_times(self._handle, _ptr(out, C.c_float), out.size, C.byref(n)))
        return out[: n.value].copy()


class _NonbondedAllPairs(Potential, _TiledExtras):
    _precision = F32

    def __init__(self, num_atoms, beta, cutoff, atom_idxs_i=None, disable_hilbert_sort=False, nblist_padding=0.1):
        h = _new_handle()
        if atom_idxs_i is None:
            idxs, n = None, -1
        else:
            idxs = _i32(atom_idxs_i)
            n = idxs.size
        _check(
            _L.tmb_nonbonded_all_pairs_create(
                self._precision, int(num_atoms), float(beta), float(cutoff), _ptr(idxs, C.c_int32), n,
                int(bool(disable_hilbert_sort)), float(nblist_padding), C.byref(h),
            )
        )
        self._adopt(h)

    def set_atom_idxs(self, atom_idxs) -> None:
        idxs = _i32(atom_idxs)
        _check(_L.tmb_nonbonded_all_pairs_set_atom_idxs(self._handle, _ptr(idxs, C.c_int32), idxs.size))

    def get_num_atom_idxs(self) -> int:
        n = C.c_int()
        _check(_L.tmb_nonbonded_all_pairs_get_num_atom_idxs(self._handle, C.byref(n)))
        return n.value

    def get_atom_idxs(self) -> list:
        out = np.empty(self.get_num_atom_idxs(), dtype=np.int32)
        _check(_L.tmb_nonbonded_all_pairs_get_atom_idxs(self._handle, _ptr(out, C.c_int32)))
        return out.tolist()

    def get_tile_count(self) -> int:
        n = C.c_uint()
        _check(_L.tmb_nonbonded_num_tiles(self._handle, C.byref(n)))
        return n.value


class NonbondedAllPairs_f32(_NonbondedAllPairs):
    _precision = F32


class NonbondedAllPairs_f64(_NonbondedAllPairs):
    _precision = F64


class _NonbondedInteractionGroup(Potential, _TiledExtras):
    _precision = F32

    def __init__(
        self, num_atoms, row_atom_idxs_i, beta, cutoff, col_atom_idxs_i=None, disable_hilbert_sort=False, nblist_padding=0.1
    ):
        rows = _i32(row_atom_idxs_i)
        if col_atom_idxs_i is None:
            cols, nc = None, -1
        else:
            cols = _i32(col_atom_idxs_i)
            nc = cols.size
        h = _new_handle()
        _check(
            _L.tmb_nonbonded_interaction_group_create(
                self._precision, int(num_atoms), _ptr(rows, C.c_int32), rows.size, float(beta), float(cutoff),
                _ptr(cols, C.c_int32), nc, int(bool(disable_hilbert_sort)), float(nblist_padding), C.byref(h),
            )
        )
        self._adopt(h)

    def set_atom_idxs(self, row_atom_idxs, col_atom_idxs) -> None:
        rows = _i32(row_atom_idxs)
        cols = _i32(col_atom_idxs)
        _check(
            _L.tmb_nonbonded_interaction_group_set_atom_idxs(
                self._handle, _ptr(rows, C.c_int32), rows.size, _ptr(cols, C.c_int32), cols.size
            )
        )

    def get_tile_count(self) -> int:
        n = C.c_uint()
        _check(_L.tmb_nonbonded_num_tiles(self._handle, C.byref(n)))
        return n.value


class NonbondedInteractionGroup_f32(_NonbondedInteractionGroup):
    _precision = F32


class NonbondedInteractionGroup_f64(_NonbondedInteractionGroup):
    _precision = F64


class _NonbondedPairList(Potential):
    _precision = F32
    _negated = False

    def __init__(self, pair_idxs_i, scales_i, beta, cutoff):
        pairs = _i32(pair_idxs_i)
        scales = _f64(scales_i)
        h = _new_handle()
        _check(
            _L.tmb_nonbonded_pair_list_create(
                self._precision, int(self._negated), _ptr(pairs, C.c_int32), pairs.size, _ptr(scales, C.c_double),
                scales.size, float(beta), float(cutoff), C.byref(h),
            )
        )
        self._adopt(h)


class NonbondedPairList_f32(_NonbondedPairList):
    _precision, _negated = F32, False


class NonbondedPairList_f64(_NonbondedPairList):
    _precision, _negated = F64, False


class NonbondedExclusions_f32(_NonbondedPairList):
    _precision, _negated = F32, True


class NonbondedExclusions_f64(_NonbondedPairList):
    _precision, _negated = F64, True


def _handle_array(objs: Sequence) -> C.Array:
    arr = (C.c_void_p * len(objs))()
    for i, o in enumerate(objs):
        arr[i] = o._handle
    return arr


class SummedPotential(Potential):
    def __init__(self, potentials, params_sizes, parallel=True):
        self._potentials = list(potentials)  # keeps the children alive, like the shared_ptrs in the reference
        sizes = _i32(params_sizes)
        h = _new_handle()
        _check(
            _L.tmb_summed_potential_create(
                _handle_array(self._potentials), len(self._potentials), _ptr(sizes, C.c_int32), sizes.size,
                int(bool(parallel)), C.byref(h),
            )
        )
        self._adopt(h)

    def get_potentials(self):
        return list(self._potentials)


class FanoutSummedPotential(Potential):
    def __init__(self, potentials, parallel=True):
        self._potentials = list(potentials)
        h = _new_handle()
        _check(
            _L.tmb_fanout_summed_potential_create(
                _handle_array(self._potentials), len(self._potentials), int(bool(parallel)), C.byref(h)
            )
        )
        self._adopt(h)

    def get_potentials(self):
        return list(self._potentials)


# ---- BoundPotential ------------------------------------------------------------------------------------------------
class BoundPotential:
    """wrap_kernels.cpp:1133-1309"""

    def __init__(self, potential: Potential, params):
        params = _f64(params)
        self._potential = potential
        self._handle = None
        h = _new_handle()
        _check(_L.tmb_bound_potential_create(potential._handle, _ptr(params, C.c_double), params.size, C.byref(h)))
        self._handle = h

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _L.tmb_bound_potential_destroy(h)
            except Exception:
                pass
            self._handle = None

    def get_potential(self) -> Potential:
        return self._potential

    def set_params(self, params) -> None:
        params = _f64(params)
        _check(_L.tmb_bound_potential_set_params(self._handle, _ptr(params, C.c_double), params.size))

    def size(self) -> int:
        n = C.c_int()
        _check(_L.tmb_bound_potential_size(self._handle, C.byref(n)))
        return n.value

    def execute(self, coords, box, compute_du_dx=True, compute_u=True):
        coords = _f64(coords)
        box = _f64(box)
        _verify_coords_and_box(coords, box)
        N = coords.shape[0]
        du_dx = np.empty((N, 3), dtype=np.uint64) if compute_du_dx else None
        u = I128() if compute_u else None
        _check(
            _L.tmb_bound_potential_execute(
                self._handle, N, _ptr(coords, C.c_double), _ptr(box, C.c_double), _ptr(du_dx, C.c_uint64),
                C.byref(u) if compute_u else None,
            )
        )
        return (_fixed_to_float(du_dx) if compute_du_dx else None, _energy_to_float(u) if compute_u else None)

    def execute_batch(self, coords, boxes, compute_du_dx, compute_u):
        coords = _f64(coords)
        boxes = _f64(boxes)
        if coords.ndim != 3 and boxes.ndim != 3:
            raise RuntimeError("coords and boxes must have 3 dimensions")
        if coords.shape[0] != boxes.shape[0]:
            raise RuntimeError("number of batches of coords and boxes don't match")
        CB, N, D = coords.shape
        du_dx = np.empty((CB, N, D), dtype=np.uint64) if compute_du_dx else None
        u = (I128 * CB)() if compute_u else None
        _check(
            _L.tmb_bound_potential_execute_batch(
                self._handle, CB, N, _ptr(coords, C.c_double), _ptr(boxes, C.c_double), _ptr(du_dx, C.c_uint64), u
            )
        )
        return (_fixed_to_float(du_dx) if compute_du_dx else None, _energies_to_float(u) if compute_u else None)

    def execute_fixed(self, coords, box) -> np.ndarray:
        coords = _f64(coords)
        box = _f64(box)
        _verify_coords_and_box(coords, box)
        u = I128()
        _check(
            _L.tmb_bound_potential_execute(
                self._handle, coords.shape[0], _ptr(coords, C.c_double), _ptr(box, C.c_double), None, C.byref(u)
            )
        )
        v = (int(u.hi) << 64) | int(u.lo)
        if v >= _LLONG_MAX or v <= _LLONG_MIN:
            v = _LLONG_MAX
        return np.array([v & ((1 << 64) - 1)], dtype=np.uint64)

    # device-pointer helpers for replica exchange drivers (BoundPotential::set_params_device, bound_potential.cu:139)
    def set_params_device(self, d_params: int, n_params: int, stream: int = 0) -> None:
        _check(_L.tmb_bound_potential_set_params_device(self._handle, C.c_void_p(d_params), n_params, C.c_void_p(stream or None)))

    def execute_device(self, N, d_coords, d_box, d_du_dx=0, d_u=0, stream=0) -> None:
        _check(
            _L.tmb_bound_potential_execute_device(
                self._handle, N, C.c_void_p(d_coords), C.c_void_p(d_box), C.c_void_p(d_du_dx or None),
                C.c_void_p(d_u or None), C.c_void_p(stream or None),
            )
        )


# ---- integrators / context -----------------------------------------------------------------------------------------
class Integrator:
    def __init__(self, *args, **kwargs):
        raise TypeError("Integrator has no constructor")


class LangevinIntegrator(Integrator):
    """LangevinIntegrator(masses, temperature, dt, friction, seed) - f32 like the reference (wrap_kernels.cpp:698-715)"""

    def __init__(self, masses, temperature, dt, friction, seed):
        masses = _f64(masses)
        self._handle = None
        h = _new_handle()
        _check(
            _L.tmb_langevin_integrator_create(
                _ptr(masses, C.c_double), masses.size, float(temperature), float(dt), float(friction), int(seed), C.byref(h)
            )
        )
        self._handle = h
        self._n = masses.size

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _L.tmb_langevin_integrator_destroy(h)
            except Exception:
                pass
            self._handle = None

    def set_noise(self, noise) -> None:
        """Test hook: use these N x 3 normals on every step (None: back to the in-kernel Philox stream)."""
        if noise is None:
            _check(_L.tmb_langevin_integrator_set_noise(self._handle, None))
            return
        noise = np.ascontiguousarray(noise, dtype=np.float32)
        if noise.shape != (self._n, 3):
            raise RuntimeError("noise must have shape (N, 3)")
        _check(_L.tmb_langevin_integrator_set_noise(self._handle, _ptr(noise, C.c_float)))

    def set_step(self, step: int) -> None:
        """Reposition the counter-based noise stream: the next step of atom a draws Philox(seed; a, step).  The
        reference's cuRAND state cannot be set (langevin_integrator.cu:35-37); hrex.py uses this to give every replica
        its own resumable sub-stream, which makes HREX trajectories independent of the replica-to-GPU layout."""
        _check(_L.tmb_langevin_integrator_set_step(self._handle, C.c_uint64(int(step) & ((1 << 63) - 1))))

    def get_step(self) -> int:
        out = C.c_uint64(0)
        _check(_L.tmb_langevin_integrator_get_step(self._handle, C.byref(out)))
        return int(out.value)


class VelocityVerletIntegrator(Integrator):
    """VelocityVerletIntegrator(dt, cbs) - f64, cbs = -dt / masses (wrap_kernels.cpp:717-729, lib/__init__.py:25-37)"""

    def __init__(self, dt, cbs):
        cbs = _f64(cbs)
        self._handle = None
        h = _new_handle()
        _check(_L.tmb_velocity_verlet_integrator_create(float(dt), _ptr(cbs, C.c_double), cbs.size, C.byref(h)))
        self._handle = h
        self._n = cbs.size

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _L.tmb_integrator_destroy(h)
            except Exception:
                pass
            self._handle = None


class Mover:
    """Base of the movers Context runs after every integrator step (wrap_kernels.cpp:1591-1617, mover.hpp:11-50)."""

    _handle = None

    def __init__(self, *args, **kwargs):
        raise TypeError("Mover: No constructor defined!")

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _L.tmb_mover_destroy(h)
            except Exception:
                pass
            self._handle = None

    def set_interval(self, interval: int) -> None:
        _check(_L.tmb_mover_set_interval(self._handle, int(interval)))

    def get_interval(self) -> int:
        n = C.c_int()
        _check(_L.tmb_mover_get_interval(self._handle, C.byref(n)))
        return n.value

    def set_step(self, step: int) -> None:
        _check(_L.tmb_mover_set_step(self._handle, int(step)))

    def move(self, coords, box):
        """One call of the mover on host arrays; returns (coords, box) after it (wrap_kernels.cpp:1599-1616)."""
        coords, box = _f64(coords), _f64(box)
        _verify_coords_and_box(coords, box)
        out_x, out_box = np.empty_like(coords), np.empty_like(box)
        _check(
            _L.tmb_mover_move_host(
                self._handle, coords.shape[0], _ptr(coords, C.c_double), _ptr(box, C.c_double), _ptr(out_x, C.c_double),
                _ptr(out_box, C.c_double),
            )
        )
        return out_x, out_box


class MonteCarloBarostat(Mover):
    """MonteCarloBarostat(N, pressure, temperature, group_idxs, interval, bps, seed, adaptive_scaling_enabled,
    initial_volume_scale_factor) (wrap_kernels.cpp:1619-1659; barostat.cu).  f32 arithmetic like the reference's only
    exported instantiation; same cuRAND stream, so the same seed gives the same accept/reject sequence."""

    def __init__(
        self, N, pressure, temperature, group_idxs, interval, bps, seed, adaptive_scaling_enabled,
        initial_volume_scale_factor,
    ):
        groups = [np.asarray(g, dtype=np.int32).reshape(-1) for g in group_idxs]
        offsets = np.zeros(len(groups) + 1, dtype=np.int32)
        if groups:
            offsets[1:] = np.cumsum([len(g) for g in groups])
        flat = np.ascontiguousarray(np.concatenate(groups) if groups else np.zeros(0, dtype=np.int32), dtype=np.int32)
        self._bps = list(bps)  # keep the Python owners alive
        self._handle = None
        h = _new_handle()
        _check(
            _L.tmb_barostat_create(
                int(N), float(pressure), float(temperature), _ptr(flat, C.c_int32), _ptr(offsets, C.c_int32), len(groups),
                int(interval), _handle_array(self._bps), len(self._bps), int(seed), int(bool(adaptive_scaling_enabled)),
                float(initial_volume_scale_factor), C.byref(h),
            )
        )
        self._handle = h

    def set_volume_scale_factor(self, volume_scale_factor: float) -> None:
        _check(_L.tmb_barostat_set_volume_scale_factor(self._handle, float(volume_scale_factor)))

    def get_volume_scale_factor(self) -> float:
        v = C.c_double()
        _check(_L.tmb_barostat_get_volume_scale_factor(self._handle, C.byref(v)))
        return v.value

    def set_adaptive_scaling(self, adaptive_scaling_enabled: bool) -> None:
        _check(_L.tmb_barostat_set_adaptive_scaling(self._handle, int(bool(adaptive_scaling_enabled))))

    def get_adaptive_scaling(self) -> bool:
        v = C.c_int()
        _check(_L.tmb_barostat_get_adaptive_scaling(self._handle, C.byref(v)))
        return bool(v.value)

    def set_pressure(self, pressure: float) -> None:
        _check(_L.tmb_barostat_set_pressure(self._handle, float(pressure)))

    # introspection, not part of the reference API
    def last_uniforms(self) -> np.ndarray:
        out = np.zeros(2, dtype=np.float32)
        _check(_L.tmb_barostat_last_uniforms(self._handle, _ptr(out, C.c_float)))
        return out

    def counters(self) -> tuple:
        out = (C.c_int * 2)()
        _check(_L.tmb_barostat_counters(self._handle, out))
        return int(out[0]), int(out[1])


def _flatten_groups(group_idxs):
    """list of index lists -> (flat int32, offsets int32, count), the layout the C ABI takes for molecules / groups."""
    groups = [np.asarray(g, dtype=np.int32).reshape(-1) for g in group_idxs]
    offsets = np.zeros(len(groups) + 1, dtype=np.int32)
    if groups:
        offsets[1:] = np.cumsum([len(g) for g in groups])
    flat = np.ascontiguousarray(np.concatenate(groups) if groups else np.zeros(0, dtype=np.int32), dtype=np.int32)
    return flat, offsets, len(groups)


class _BDExchangeMove(Mover):
    """BDExchangeMove_{f32,f64}(N, target_mols, params, temperature, nb_beta, cutoff, seed, num_proposals_per_move,
    interval, batch_size=1): water exchange by biased deletion (wrap_kernels.cpp:1732-1900; bd_exchange_move.cu; the
    Python reference is timemachine/md/exchange/exchange_mover.py::BDExchangeMove).  Same four cuRAND streams as the
    reference: the same seed proposes and accepts the same moves."""

    _precision = F32

    def __init__(
        self, N, target_mols, params, temperature, nb_beta, cutoff, seed, num_proposals_per_move, interval, batch_size=1
    ):
        params = np.asarray(params, dtype=np.float64)
        if int(num_proposals_per_move) <= 0:
            raise RuntimeError("proposals per move must be greater than 0")
        if params.ndim != 2:
            raise RuntimeError("parameters dimensions must be 2")
        if params.shape[0] != int(N):
            raise RuntimeError("Number of parameters must match N")
        params = np.ascontiguousarray(params)
        flat, offsets, n_mols = _flatten_groups(target_mols)
        self._handle = None
        h = _new_handle()
        _check(
            _L.tmb_bd_exchange_move_create(
                self._precision, int(N), _ptr(flat, C.c_int32), _ptr(offsets, C.c_int32), n_mols, _ptr(params, C.c_double),
                int(params.size), float(temperature), float(nb_beta), float(cutoff), int(seed), int(num_proposals_per_move),
                int(interval), int(batch_size), C.byref(h),
            )
        )
        self._handle = h
        self._N = int(N)
        self._real = np.float32 if self._precision == F32 else np.float64

    def _num_target_mols(self) -> int:
        n = C.c_int()
        _check(_L.tmb_bd_exchange_move_num_target_mols(self._handle, C.byref(n)))
        return n.value

    def batch_size(self) -> int:
        n = C.c_int()
        _check(_L.tmb_bd_exchange_move_batch_size(self._handle, C.byref(n)))
        return n.value

    def compute_initial_log_weights(self, coords, box) -> list:
        coords, box = _f64(coords), _f64(box)
        _verify_coords_and_box(coords, box)
        out = np.empty(self._num_target_mols(), dtype=np.float64)
        _check(
            _L.tmb_bd_exchange_move_initial_log_weights(
                self._handle, coords.shape[0], _ptr(coords, C.c_double), _ptr(box, C.c_double), _ptr(out, C.c_double)
            )
        )
        return list(out.astype(self._real))

    def compute_incremental_log_weights(self, coords, box, mol_idxs, quaternions, translation) -> list:
        coords, box = _f64(coords), _f64(box)
        _verify_coords_and_box(coords, box)
        mol_idxs = _i32(mol_idxs).reshape(-1)
        quaternions = _f64(quaternions)
        translation = _f64(translation)
        B = self.batch_size()
        if mol_idxs.size != B:
            raise RuntimeError("number of mol idxs must match batch size")
        if quaternions.shape[0] != B:
            raise RuntimeError("number of quaternions must match batch size")
        if quaternions.shape[1] != 4:
            raise RuntimeError("each quaternion must be of length 4")
        if translation.shape[0] != B:
            raise RuntimeError("number of translations must match batch size")
        if translation.shape[1] != 3:
            raise RuntimeError("each translation must be of length 3")
        M = self._num_target_mols()
        out = np.empty((B, M), dtype=np.float64)
        _check(
            _L.tmb_bd_exchange_move_incremental_log_weights(
                self._handle, coords.shape[0], _ptr(coords, C.c_double), _ptr(box, C.c_double), _ptr(mol_idxs, C.c_int32),
                _ptr(quaternions, C.c_double), _ptr(translation, C.c_double), _ptr(out, C.c_double),
            )
        )
        return [list(row.astype(self._real)) for row in out]

    def get_params(self) -> np.ndarray:
        out = np.empty((self._N, 4), dtype=np.float64)
        _check(_L.tmb_bd_exchange_move_get_params(self._handle, _ptr(out, C.c_double), int(out.size)))
        return out

    def set_params(self, params) -> None:
        params = _f64(params)
        _check(_L.tmb_bd_exchange_move_set_params(self._handle, _ptr(params, C.c_double), int(params.size)))

    def last_log_probability(self) -> float:
        v = C.c_double()
        _check(_L.tmb_bd_exchange_move_last_log_probability(self._handle, C.byref(v)))
        return v.value

    def last_raw_log_probability(self) -> float:
        v = C.c_double()
        _check(_L.tmb_bd_exchange_move_last_raw_log_probability(self._handle, C.byref(v)))
        return v.value

    def n_accepted(self) -> int:
        v = C.c_ulonglong()
        _check(_L.tmb_bd_exchange_move_n_accepted(self._handle, C.byref(v)))
        return int(v.value)

    def n_proposed(self) -> int:
        v = C.c_ulonglong()
        _check(_L.tmb_bd_exchange_move_n_proposed(self._handle, C.byref(v)))
        return int(v.value)

    def acceptance_fraction(self) -> float:
        return self.n_accepted() / self.n_proposed()

    def get_before_log_weights(self) -> list:
        out = np.empty(self._num_target_mols(), dtype=np.float64)
        _check(_L.tmb_bd_exchange_move_before_log_weights(self._handle, _ptr(out, C.c_double)))
        return list(out.astype(self._real))

    def get_after_log_weights(self) -> list:
        out = np.empty(self._num_target_mols() * self.batch_size(), dtype=np.float64)
        _check(_L.tmb_bd_exchange_move_after_log_weights(self._handle, _ptr(out, C.c_double)))
        return list(out.astype(self._real))


class BDExchangeMove_f32(_BDExchangeMove):
    _precision = F32


class BDExchangeMove_f64(_BDExchangeMove):
    _precision = F64


class _TIBDExchangeMove:
    """TIBDExchangeMove_{f32,f64}(N, ligand_idxs, target_mols, params, temperature, nb_beta, cutoff, radius, seed,
    num_proposals_per_move, interval, batch_size=1): targeted insertion / biased deletion between a sphere around the
    ligand centroid and the bulk (wrap_kernels.cpp:1902-1975; tibd_exchange_move.cu; the Python reference is
    timemachine/md/exchange/exchange_mover.py::TIBDExchangeMove)."""

    def __init__(
        self, N, ligand_idxs, target_mols, params, temperature, nb_beta, cutoff, radius, seed, num_proposals_per_move,
        interval, batch_size=1,
    ):
        params = np.asarray(params, dtype=np.float64)
        if int(num_proposals_per_move) <= 0:
            raise RuntimeError("proposals per move must be greater than 0")
        if params.ndim != 2:
            raise RuntimeError("parameters dimensions must be 2")
        if params.shape[0] != int(N):
            raise RuntimeError("Number of parameters must match N")
        params = np.ascontiguousarray(params)
        ligand_idxs = _i32(ligand_idxs).reshape(-1)
        flat, offsets, n_mols = _flatten_groups(target_mols)
        self._handle = None
        h = _new_handle()
        _check(
            _L.tmb_tibd_exchange_move_create(
                self._precision, int(N), _ptr(ligand_idxs, C.c_int32), int(ligand_idxs.size), _ptr(flat, C.c_int32),
                _ptr(offsets, C.c_int32), n_mols, _ptr(params, C.c_double), int(params.size), float(temperature), float(nb_beta),
                float(cutoff), float(radius), int(seed), int(num_proposals_per_move), int(interval), int(batch_size), C.byref(h),
            )
        )
        self._handle = h
        self._N = int(N)
        self._real = np.float32 if self._precision == F32 else np.float64


class TIBDExchangeMove_f32(_TIBDExchangeMove, BDExchangeMove_f32):  # custom_ops.pyi: TIBDExchangeMove_f32(BDExchangeMove_f32)
    _precision = F32


class TIBDExchangeMove_f64(_TIBDExchangeMove, BDExchangeMove_f64):
    _precision = F64


def _inner_and_outer_mols(precision, center_atoms, coords, box, group_idxs, radius):
    coords, box = _f64(coords), _f64(box)
    _verify_coords_and_box(coords, box)
    center_atoms = _i32(center_atoms).reshape(-1)
    flat, offsets, n_mols = _flatten_groups(group_idxs)
    flags = np.zeros(n_mols, dtype=np.int32)
    _check(
        _L.tmb_inner_and_outer_mols(
            precision, _ptr(center_atoms, C.c_int32), int(center_atoms.size), coords.shape[0], _ptr(coords, C.c_double),
            _ptr(box, C.c_double), _ptr(flat, C.c_int32), _ptr(offsets, C.c_int32), n_mols, float(radius), _ptr(flags, C.c_int32),
        )
    )
    idx = np.arange(n_mols)
    return [int(i) for i in idx[flags == 1]], [int(i) for i in idx[flags != 1]]


def inner_and_outer_mols_f32(center_atoms, coords, box, group_idxs, radius):
    """(inner, outer) molecule indices: centroid within `radius` of the centroid of center_atoms, minimum image
    (wrap_kernels.cpp:2032-2048)."""
    return _inner_and_outer_mols(F32, center_atoms, coords, box, group_idxs, radius)


def inner_and_outer_mols_f64(center_atoms, coords, box, group_idxs, radius):
    return _inner_and_outer_mols(F64, center_atoms, coords, box, group_idxs, radius)


def _translations_inside_and_outside_sphere_host(precision, num_translations, box, center, radius, seed):
    box, center = _f64(box), _f64(center).reshape(-1)
    if center.size != 3:
        raise RuntimeError("Center must be of length 3")
    out = np.empty((int(num_translations), 2, 3), dtype=np.float64)
    _check(
        _L.tmb_translations_inside_and_outside_sphere(
            precision, int(num_translations), _ptr(box, C.c_double), _ptr(center, C.c_double), float(radius), int(seed),
            _ptr(out, C.c_double),
        )
    )
    return out.astype(np.float32 if precision == F32 else np.float64)


def translations_inside_and_outside_sphere_host_f32(num_translations, box, center, radius, seed):
    """[n, 2, 3]: a translation inside the sphere around `center` and one outside it, per row (wrap_kernels.cpp:2117-2141)."""
    return _translations_inside_and_outside_sphere_host(F32, num_translations, box, center, radius, seed)


def translations_inside_and_outside_sphere_host_f64(num_translations, box, center, radius, seed):
    return _translations_inside_and_outside_sphere_host(F64, num_translations, box, center, radius, seed)


class _NonbondedMolEnergyPotential:
    """NonbondedMolEnergyPotential_{f32,f64}(N, target_mols, beta, cutoff).execute(coords, params, box) -> energy of
    every target molecule with all atoms outside it (wrap_kernels.cpp:234-294; nonbonded_mol_energy.cu)."""

    _precision = F32

    def __init__(self, N, target_mols, beta, cutoff):
        self._N, self._beta, self._cutoff = int(N), float(beta), float(cutoff)
        self._flat, self._offsets, self._n_mols = _flatten_groups(target_mols)
        self._call(None, None, None, None)  # the constructor's argument checks

    def _call(self, coords, params, box, out):
        _check(
            _L.tmb_nonbonded_mol_energies(
                self._precision, self._N, _ptr(self._flat, C.c_int32), _ptr(self._offsets, C.c_int32), self._n_mols,
                self._beta, self._cutoff, _ptr(coords, C.c_double), _ptr(params, C.c_double), _ptr(box, C.c_double), out,
            )
        )

    def execute(self, coords, params, box) -> np.ndarray:
        coords, params, box = _f64(coords), _f64(params), _f64(box)
        _verify_coords_and_box(coords, box)
        if coords.shape[0] != params.shape[0]:
            raise RuntimeError("params N != coords N")
        out = (I128 * self._n_mols)()
        self._call(coords, params, box, out)
        return _energies_to_float(out)


class NonbondedMolEnergyPotential_f32(_NonbondedMolEnergyPotential):
    _precision = F32


class NonbondedMolEnergyPotential_f64(_NonbondedMolEnergyPotential):
    _precision = F64


def _atom_by_atom_energies(precision, target_atoms, coords, params, box, nb_beta, cutoff):
    coords, params, box = _f64(coords), _f64(params), _f64(box)
    _verify_coords_and_box(coords, box)
    target_atoms = _i32(target_atoms).reshape(-1)
    out = np.empty((target_atoms.size, coords.shape[0]), dtype=np.float64)
    _check(
        _L.tmb_atom_by_atom_energies(
            precision, coords.shape[0], _ptr(target_atoms, C.c_int32), int(target_atoms.size), _ptr(coords, C.c_double),
            _ptr(params, C.c_double), _ptr(box, C.c_double), float(nb_beta), float(cutoff), _ptr(out, C.c_double),
        )
    )
    return out.astype(np.float32 if precision == F32 else np.float64)


def atom_by_atom_energies_f32(target_atoms, coords, params, box, nb_beta, nb_cutoff):
    """Pair energies of the target atoms with every atom, [T, N] (wrap_kernels.cpp:2004-2030)."""
    return _atom_by_atom_energies(F32, target_atoms, coords, params, box, nb_beta, nb_cutoff)


def atom_by_atom_energies_f64(target_atoms, coords, params, box, nb_beta, nb_cutoff):
    return _atom_by_atom_energies(F64, target_atoms, coords, params, box, nb_beta, nb_cutoff)


def _flatten_values(values):
    rows = [np.asarray(v, dtype=np.float64).reshape(-1) for v in values]
    offsets = np.zeros(len(rows) + 1, dtype=np.int32)
    if rows:
        offsets[1:] = np.cumsum([len(r) for r in rows])
    flat = np.ascontiguousarray(np.concatenate(rows) if rows else np.zeros(0), dtype=np.float64)
    return flat, offsets, len(rows)


class _SegmentedSumExp:
    """SegmentedSumExp_{f32,f64}(max_vals_per_segment, num_segments).logsumexp(values) (wrap_kernels.cpp:1693-1730)."""

    _precision = F32

    def __init__(self, max_vals_per_segment, num_segments):
        self._max_vals, self._num_segments = int(max_vals_per_segment), int(num_segments)

    def logsumexp(self, values) -> list:
        flat, offsets, n = _flatten_values(values)
        out = np.empty(n, dtype=np.float64)
        _check(
            _L.tmb_segmented_logsumexp(
                self._precision, self._max_vals, self._num_segments, _ptr(flat, C.c_double), _ptr(offsets, C.c_int32), n,
                _ptr(out, C.c_double),
            )
        )
        return list(out.astype(np.float32 if self._precision == F32 else np.float64))


class SegmentedSumExp_f32(_SegmentedSumExp):
    _precision = F32


class SegmentedSumExp_f64(_SegmentedSumExp):
    _precision = F64


class _SegmentedWeightedRandomSampler:
    """SegmentedWeightedRandomSampler_{f32,f64}(max_vals_per_segment, segments, seed).sample(weights): one index per
    segment, drawn with probability proportional to its weight (Gumbel-max trick; wrap_kernels.cpp:196-232)."""

    _precision = F32
    _handle = None

    def __init__(self, max_vals_per_segment, segments, seed):
        h = _new_handle()
        _check(_L.tmb_weighted_sampler_create(self._precision, int(max_vals_per_segment), int(segments), int(seed), C.byref(h)))
        self._handle = h

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _L.tmb_weighted_sampler_destroy(h)
            except Exception:
                pass
            self._handle = None

    def sample(self, weights) -> list:
        flat, offsets, n = _flatten_values(weights)
        out = np.empty(n, dtype=np.int32)
        _check(_L.tmb_weighted_sampler_sample(self._handle, _ptr(flat, C.c_double), _ptr(offsets, C.c_int32), n, _ptr(out, C.c_int32)))
        return [int(v) for v in out]


class SegmentedWeightedRandomSampler_f32(_SegmentedWeightedRandomSampler):
    _precision = F32


class SegmentedWeightedRandomSampler_f64(_SegmentedWeightedRandomSampler):
    _precision = F64


def _rotate_coords(precision, coords, quaternions):
    coords, quaternions = _f64(coords), _f64(quaternions)
    _verify_coords(coords)
    if quaternions.ndim != 2:
        raise RuntimeError("quaternions dimensions must be 2")
    if quaternions.shape[-1] != 4:
        raise RuntimeError("quaternions must have a shape that is 4 dimensional")
    out = np.empty((coords.shape[0], quaternions.shape[0], 3), dtype=np.float64)
    _check(
        _L.tmb_rotate_coords(
            precision, coords.shape[0], quaternions.shape[0], _ptr(coords, C.c_double), _ptr(quaternions, C.c_double),
            _ptr(out, C.c_double),
        )
    )
    return out


def rotate_coords_f32(coords, quaternions):
    """Rotate every coordinate by every quaternion, [N, R, 3] (wrap_kernels.cpp:2050-2070)."""
    return _rotate_coords(F32, coords, quaternions)


def rotate_coords_f64(coords, quaternions):
    return _rotate_coords(F64, coords, quaternions)


def _rotate_and_translate_mol(precision, coords, box, quaternions, translations):
    coords, box, quaternions, translations = _f64(coords), _f64(box), _f64(quaternions), _f64(translations)
    _verify_coords_and_box(coords, box)
    if quaternions.ndim != 2:
        raise RuntimeError("quaternions dimensions must be 2")
    if quaternions.shape[1] != 4:
        raise RuntimeError("quaternions must be of length 4")
    if translations.ndim != 2:
        raise RuntimeError("translations dimensions must be 2")
    if translations.shape[1] != 3:
        raise RuntimeError("translations must be of size 3")
    if quaternions.shape[0] != translations.shape[0]:
        raise RuntimeError("Number of quaternions and translations must match")
    out = np.empty((quaternions.shape[0], coords.shape[0], 3), dtype=np.float64)
    _check(
        _L.tmb_rotate_and_translate_mol(
            precision, coords.shape[0], quaternions.shape[0], _ptr(coords, C.c_double), _ptr(box, C.c_double),
            _ptr(quaternions, C.c_double), _ptr(translations, C.c_double), _ptr(out, C.c_double),
        )
    )
    return out


def rotate_and_translate_mol_f32(coords, box, quaternion, translation):
    """Rotate a molecule about its centroid and move the centroid to translation * box, per (quaternion, translation)
    pair, [B, N, 3] (wrap_kernels.cpp:2072-2115; the keywords are singular there although both are batches)."""
    return _rotate_and_translate_mol(F32, coords, box, quaternion, translation)


def rotate_and_translate_mol_f64(coords, box, quaternion, translation):
    return _rotate_and_translate_mol(F64, coords, box, quaternion, translation)


class Context:
    """Context(x0, v0, box, integrator, bps, movers=None) (wrap_kernels.cpp:296-689).  Movers run after every
    integrator step (context.cu:261-277); MonteCarloBarostat is the one implemented here."""

    def __init__(self, x0, v0, box, integrator, bps, movers=None):
        x0 = _f64(x0)
        v0 = _f64(v0)
        box = _f64(box)
        _verify_coords_and_box(x0, box)
        if v0.ndim != 2 or x0.shape[0] != v0.shape[0]:
            raise RuntimeError("v0 N != x0 N")
        if x0.shape[1] != v0.shape[1]:
            raise RuntimeError("v0 D != x0 D")
        self._movers = list(movers) if movers else []
        for m in self._movers:
            if not isinstance(m, Mover):
                raise RuntimeError("movers must be timemachine_b200 Mover objects")
        if not isinstance(integrator, Integrator) or getattr(integrator, "_handle", None) is None:
            raise RuntimeError("integrator must be a timemachine_b200 Integrator object")
        self._integrator = integrator
        self._bps = list(bps)
        self._n = x0.shape[0]
        self._handle = None
        h = _new_handle()
        _check(
            _L.tmb_context_create_with_movers(
                _ptr(x0, C.c_double), _ptr(v0, C.c_double), _ptr(box, C.c_double), self._n, integrator._handle,
                _handle_array(self._bps), len(self._bps), _handle_array(self._movers), len(self._movers), C.byref(h),
            )
        )
        self._handle = h

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _L.tmb_context_destroy(h)
            except Exception:
                pass
            self._handle = None

    def step(self) -> None:
        _check(_L.tmb_context_step(self._handle))

    def initialize(self) -> None:
        """The integrator's opening half step (context.cu:256-260): nothing for Langevin, half kick + drift for Verlet."""
        _check(_L.tmb_context_initialize(self._handle))

    def finalize(self) -> None:
        _check(_L.tmb_context_finalize(self._handle))

    def multiple_steps(self, n_steps: int, store_x_interval: int = 0):
        if store_x_interval < 0:
            raise RuntimeError("store_x_interval must be greater than or equal to zero")
        n_steps = int(n_steps)
        x_interval = n_steps if store_x_interval == 0 else int(store_x_interval)
        n_samples = n_steps // x_interval if x_interval > 0 else 0
        xs = np.empty((n_samples, self._n, 3), dtype=np.float64)
        boxes = np.empty((n_samples, 3, 3), dtype=np.float64)
        _check(_L.tmb_context_multiple_steps(self._handle, n_steps, n_samples, _ptr(xs, C.c_double), _ptr(boxes, C.c_double)))
        return xs, boxes

    # ---- local MD (wrap_kernels.cpp:368-612): a shell of atoms around one frozen reference atom is simulated -------------
    @staticmethod
    def _verify_local_md_parameters(radius: float, k: float) -> None:
        """local_md_utils.cu:95-116"""
        if radius < 0.1:
            raise RuntimeError("radius must be greater or equal to 0.100000")
        if k < 1.0:
            raise RuntimeError("k must be at least one")
        if k > 1e6:
            raise RuntimeError("k must be less than than 1e+06")

    def _verify_atom_idxs(self, idxs: np.ndarray) -> None:
        """verify_atom_idxs (nonbonded_common.cpp): same checks and messages as the potentials' index arguments"""
        if idxs.size == 0:
            raise RuntimeError("indices can't be empty")
        if np.unique(idxs).size != idxs.size:
            raise RuntimeError("atom indices must be unique")
        if idxs.max() >= self._n:
            raise RuntimeError(f"index values must be less than N({self._n})")
        if idxs.min() < 0:
            raise RuntimeError("index values must be greater or equal to zero")

    def setup_local_md(self, temperature: float, freeze_reference: bool) -> None:
        """Configure local MD before its first use (default on first use: the integrator's temperature, frozen reference)."""
        _check(_L.tmb_context_setup_local_md(self._handle, float(temperature), int(bool(freeze_reference))))

    def _local_frames(self, n_steps: int, store_x_interval: int):
        if n_steps <= 0:
            raise RuntimeError("local steps must be at least one")
        if store_x_interval < 0:
            raise RuntimeError("store_x_interval must be greater than or equal to zero")
        x_interval = n_steps if store_x_interval == 0 else int(store_x_interval)
        n_samples = n_steps // x_interval
        return n_samples, np.empty((n_samples, self._n, 3), dtype=np.float64), np.empty((n_samples, 3, 3), dtype=np.float64)

    def multiple_steps_local(self, n_steps: int, local_idxs, store_x_interval: int = 0, radius: float = 1.2, k: float = 10000.0, seed: int = 2022):
        """Steps of the atoms selected (with probability exp(-U_flat_bottom / kT)) around one atom drawn from local_idxs;
        that atom and everything unselected stay frozen.  Movers (barostat) do not run.  Returns (xs, boxes)."""
        n_steps = int(n_steps)
        n_samples, xs, boxes = self._local_frames(n_steps, int(store_x_interval))
        self._verify_local_md_parameters(float(radius), float(k))
        idxs = _i32(local_idxs).reshape(-1)
        self._verify_atom_idxs(idxs)
        _check(
            _L.tmb_context_multiple_steps_local(
                self._handle, n_steps, _ptr(idxs, C.c_int32), idxs.size, n_samples, float(radius), float(k), int(seed),
                _ptr(xs, C.c_double), _ptr(boxes, C.c_double),
            )
        )
        return xs, boxes

    def multiple_steps_local_selection(self, n_steps: int, reference_idx: int, selection_idxs, store_x_interval: int = 0, radius: float = 1.2, k: float = 10000.0):
        """Steps of the given selection of free atoms, restrained to reference_idx by a flat-bottom bond."""
        n_steps = int(n_steps)
        n_samples, xs, boxes = self._local_frames(n_steps, int(store_x_interval))
        self._verify_local_md_parameters(float(radius), float(k))
        reference_idx = int(reference_idx)
        if reference_idx < 0 or reference_idx >= self._n:
            raise RuntimeError(f"reference idx must be at least 0 and less than {self._n}")
        idxs = _i32(selection_idxs).reshape(-1)
        self._verify_atom_idxs(idxs)
        if reference_idx in set(idxs.tolist()):
            raise RuntimeError("reference idx must not be in selection idxs")
        _check(
            _L.tmb_context_multiple_steps_local_selection(
                self._handle, n_steps, reference_idx, _ptr(idxs, C.c_int32), idxs.size, n_samples, float(radius), float(k),
                _ptr(xs, C.c_double), _ptr(boxes, C.c_double),
            )
        )
        return xs, boxes

    def local_md_free_idxs(self) -> np.ndarray:
        """Atoms that were free in the last local MD call (introspection; not part of the reference API)."""
        out = np.empty(self._n, dtype=np.uint32)
        _check(_L.tmb_context_local_md_free_idxs(self._handle, _ptr(out, C.c_uint32)))
        return np.flatnonzero(out < self._n).astype(np.int32)

    def set_x_t(self, coords) -> None:
        coords = _f64(coords)
        if coords.shape != (self._n, 3):
            raise RuntimeError("number of new coords disagree with current coords")
        _check(_L.tmb_context_set_x_t(self._handle, _ptr(coords, C.c_double)))

    def set_v_t(self, velocities) -> None:
        velocities = _f64(velocities)
        if velocities.shape != (self._n, 3):
            raise RuntimeError("number of new velocities disagree with current coords")
        _check(_L.tmb_context_set_v_t(self._handle, _ptr(velocities, C.c_double)))

    def set_box(self, box) -> None:
        box = _f64(box)
        if box.shape != (3, 3):
            raise RuntimeError("box must be 3x3")
        _check(_L.tmb_context_set_box(self._handle, _ptr(box, C.c_double)))

    def get_x_t(self) -> np.ndarray:
        out = np.empty((self._n, 3), dtype=np.float64)
        _check(_L.tmb_context_get_x_t(self._handle, _ptr(out, C.c_double)))
        return out

    def get_v_t(self) -> np.ndarray:
        out = np.empty((self._n, 3), dtype=np.float64)
        _check(_L.tmb_context_get_v_t(self._handle, _ptr(out, C.c_double)))
        return out

    def get_box(self) -> np.ndarray:
        out = np.empty((3, 3), dtype=np.float64)
        _check(_L.tmb_context_get_box(self._handle, _ptr(out, C.c_double)))
        return out

    def get_integrator(self):
        return self._integrator

    def get_potentials(self):
        return list(self._bps)

    def get_movers(self):
        return list(self._movers)

    def get_barostat(self):
        """The first MonteCarloBarostat among the movers, else None (context.cu:311-320)."""
        for m in self._movers:
            if isinstance(m, MonteCarloBarostat):
                return m
        return None

    # extensions used by bench.py / the replica driver
    def set_stream(self, stream: int) -> None:
        _check(_L.tmb_context_set_stream(self._handle, C.c_void_p(stream or None)))

    def set_use_graphs(self, on: bool) -> None:
        _check(_L.tmb_context_set_use_graphs(self._handle, int(bool(on))))

    def device_state(self):
        """(d_x, d_v, d_box) raw device pointers of the context's state."""
        dx, dv, db = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(_L.tmb_context_device_state(self._handle, C.byref(dx), C.byref(dv), C.byref(db)))
        return dx.value, dv.value, db.value


# ---- neighbour list / Hilbert sort ------------------------------------------------------------------------------------
def rmsd_align(x1, x2) -> np.ndarray:
    """x2 rotated (no reflection) and shifted onto x1 so that the RMSD is minimal (wrap_kernels.cpp:1975-2001)."""
    x1 = _f64(x1)
    x2 = _f64(x2)
    if x1.shape[0] != x2.shape[0]:
        raise RuntimeError("N1 != N2")
    if x1.ndim != 2 or x1.shape[1] != 3:
        raise RuntimeError("D1 != 3")
    if x2.ndim != 2 or x2.shape[1] != 3:
        raise RuntimeError("D2 != 3")
    out = np.empty_like(x1)
    _check(_L.tmb_rmsd_align(_ptr(x1, C.c_double), _ptr(x2, C.c_double), x1.shape[0], _ptr(out, C.c_double)))
    return out


class _Neighborlist:
    _precision = F32

    def __init__(self, N: int):
        self._handle = None
        h = _new_handle()
        _check(_L.tmb_neighborlist_create(self._precision, int(N), C.byref(h)))
        self._handle = h

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _L.tmb_neighborlist_destroy(h)
            except Exception:
                pass
            self._handle = None

    def get_nblist(self, coords, box, cutoff):
        coords = _f64(coords)
        box = _f64(box)
        _verify_coords_and_box(coords, box)
        nrows, nent = C.c_int(), C.c_int()
        _check(
            _L.tmb_neighborlist_build(
                self._handle, coords.shape[0], _ptr(coords, C.c_double), _ptr(box, C.c_double), float(cutoff),
                C.byref(nrows), C.byref(nent),
            )
        )
        offsets = np.empty(nrows.value + 1, dtype=np.int32)
        atoms = np.empty(max(nent.value, 1), dtype=np.int32)
        _check(_L.tmb_neighborlist_fetch(self._handle, _ptr(offsets, C.c_int32), _ptr(atoms, C.c_int32)))
        return [atoms[offsets[r] : offsets[r + 1]].tolist() for r in range(nrows.value)]

    def compute_block_bounds(self, coords, box, block_size):
        if block_size != 32:
            raise RuntimeError("Block size must be 32.")
        coords = _f64(coords)
        box = _f64(box)
        _verify_coords_and_box(coords, box)
        N = coords.shape[0]
        B = (N + block_size - 1) // block_size
        ctrs = np.empty((B, 3), dtype=np.float64)
        exts = np.empty((B, 3), dtype=np.float64)
        _check(
            _L.tmb_neighborlist_compute_block_bounds(
                self._handle, N, _ptr(coords, C.c_double), _ptr(box, C.c_double), _ptr(ctrs, C.c_double), _ptr(exts, C.c_double)
            )
        )
        return ctrs, exts

    def set_row_idxs(self, idxs) -> None:
        idxs = _u32(idxs)
        _check(_L.tmb_neighborlist_set_row_idxs(self._handle, _ptr(idxs, C.c_uint32), idxs.size))

    def reset_row_idxs(self) -> None:
        _check(_L.tmb_neighborlist_reset_row_idxs(self._handle))

    def resize(self, size: int) -> None:
        _check(_L.tmb_neighborlist_resize(self._handle, int(size)))

    def get_tile_ixn_count(self) -> int:
        n = C.c_uint()
        _check(_L.tmb_neighborlist_get_tile_ixn_count(self._handle, C.byref(n)))
        return n.value

    def get_max_ixn_count(self) -> int:
        n = C.c_int()
        _check(_L.tmb_neighborlist_get_max_ixn_count(self._handle, C.byref(n)))
        return n.value

    def get_num_row_idxs(self) -> int:
        n = C.c_int()
        _check(_L.tmb_neighborlist_get_num_row_idxs(self._handle, C.byref(n)))
        return n.value


class Neighborlist_f32(_Neighborlist):
    _precision = F32


class Neighborlist_f64(_Neighborlist):
    _precision = F64


class HilbertSort:
    def __init__(self, size: int):
        self._handle = None
        h = _new_handle()
        _check(_L.tmb_hilbert_sort_create(int(size), C.byref(h)))
        self._handle = h

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _L.tmb_hilbert_sort_destroy(h)
            except Exception:
                pass
            self._handle = None

    def sort(self, coords, box) -> np.ndarray:
        coords = _f64(coords)
        box = _f64(box)
        _verify_coords_and_box(coords, box)
        perm = np.empty(coords.shape[0], dtype=np.uint32)
        _check(
            _L.tmb_hilbert_sort_sort(
                self._handle, coords.shape[0], _ptr(coords, C.c_double), _ptr(box, C.c_double), _ptr(perm, C.c_uint32)
            )
        )
        return perm


def fill_normal(n_atoms: int, seed: int, step: int) -> np.ndarray:
    """N x 3 standard normals from the integrator's Philox stream (test utility)."""
    out = np.empty((n_atoms, 3), dtype=np.float32)
    _check(_L.tmb_fill_normal(_ptr(out, C.c_float), int(n_atoms), int(seed), int(step)))
    return out
