"""Build the C restatement of the oracle (test infrastructure) into oracle/_build/libtm_oracle.so with gcc + OpenMP."""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "tm_oracle_c.c"
OUT = HERE / "_build" / "libtm_oracle.so"


def build(force: bool = False) -> Path:
    OUT.parent.mkdir(exist_ok=True)
    if force or not OUT.exists() or OUT.stat().st_mtime < SRC.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", str(OUT), str(SRC), "-lm"])
    return OUT


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        d, i = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _lib.tmo_nonbonded_block.restype = C.c_double
        _lib.tmo_nonbonded_block.argtypes = [C.c_int, d, d, d, i, C.c_int, i, C.c_int, C.c_int, C.c_double, C.c_double, d, d]
        _lib.tmo_nonbonded_pairs.restype = C.c_double
        _lib.tmo_nonbonded_pairs.argtypes = [C.c_int, d, d, d, i, d, C.c_int, C.c_double, C.c_double, C.c_double, d, d]
        _lib.tmo_harmonic_bond.restype = C.c_double
        _lib.tmo_harmonic_bond.argtypes = [C.c_int, d, d, i, d]
        _lib.tmo_harmonic_angle.restype = C.c_double
        _lib.tmo_harmonic_angle.argtypes = [C.c_int, d, d, i, d]
        _lib.tmo_baoab.restype = None
        _lib.tmo_baoab.argtypes = [C.c_int, d, d, d, C.c_double, d, d, C.c_double, d]
        _lib.tmo_num_threads.restype = C.c_int
        _lib.tmo_set_num_threads.restype = None
        _lib.tmo_set_num_threads.argtypes = [C.c_int]
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def num_threads() -> int:
    return int(lib().tmo_num_threads())


def set_num_threads(n: int) -> None:
    """omp_set_num_threads: overrides an inherited OMP_NUM_THREADS (torchrun sets it to 1 for its workers)."""
    lib().tmo_set_num_threads(int(n))


def nonbonded_block(x, params, box, rows, cols, beta, cutoff, triangular, want_dx=True, want_dp=False):
    L = lib()
    x, px = _d(x)
    params, pp = _d(params)
    box, pb = _d(box)
    rows, pr = _i(rows)
    cols, pc = _i(cols)
    N = len(x)
    dx = np.zeros((N, 3)) if want_dx else None
    dp = np.zeros((N, 4)) if want_dp else None
    u = L.tmo_nonbonded_block(
        N, px, pp, pb, pr, len(rows), pc, len(cols), int(triangular), float(beta), float(cutoff),
        dx.ctypes.data_as(C.POINTER(C.c_double)) if want_dx else None,
        dp.ctypes.data_as(C.POINTER(C.c_double)) if want_dp else None,
    )
    return u, dx, dp


def nonbonded_pairs(x, params, box, pairs, scales, sign, beta, cutoff, dx=None, dp=None):
    L = lib()
    x, px = _d(x)
    params, pp = _d(params)
    box, pb = _d(box)
    pairs, ppairs = _i(np.asarray(pairs).reshape(-1, 2))
    scales, ps = _d(np.asarray(scales).reshape(-1, 2))
    return L.tmo_nonbonded_pairs(
        len(x), px, pp, pb, ppairs, ps, len(pairs), float(sign), float(beta), float(cutoff),
        dx.ctypes.data_as(C.POINTER(C.c_double)) if dx is not None else None,
        dp.ctypes.data_as(C.POINTER(C.c_double)) if dp is not None else None,
    )


def harmonic_bond(x, p, idxs, dx=None):
    L = lib()
    x, px = _d(x)
    p, pp = _d(p)
    idxs, pi = _i(idxs)
    return L.tmo_harmonic_bond(len(idxs), px, pp, pi, dx.ctypes.data_as(C.POINTER(C.c_double)) if dx is not None else None)


def harmonic_angle(x, p, idxs, dx=None):
    L = lib()
    x, px = _d(x)
    p, pp = _d(p)
    idxs, pi = _i(idxs)
    return L.tmo_harmonic_angle(len(idxs), px, pp, pi, dx.ctypes.data_as(C.POINTER(C.c_double)) if dx is not None else None)


def baoab(x, v, du_dx, ca, cb, cc, dt, noise):
    """In-place float64 BAOAB step on contiguous arrays."""
    L = lib()
    assert x.flags.c_contiguous and v.flags.c_contiguous and x.dtype == np.float64 and v.dtype == np.float64
    _, pdu = _d(du_dx)
    cb, pcb = _d(cb)
    cc, pcc = _d(cc)
    noise, pn = _d(noise)
    L.tmo_baoab(len(x), x.ctypes.data_as(C.POINTER(C.c_double)), v.ctypes.data_as(C.POINTER(C.c_double)), pdu, float(ca), pcb, pcc, float(dt), pn)


if __name__ == "__main__":
    print(build(force=True))
