"""Hilbert curve: the product's index function (csrc/hilbert_curve.h, compiled here for the host), the oracle's numpy
restatement and the reference's vendored C routine must agree on all 128^3 cells; GPU: the device LUT + sort give the
permutation the oracle predicts, bit for bit."""

import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import tm_oracle as O

ROOT = Path(__file__).resolve().parents[1]
REF_LIB = ROOT / "oracle" / "_ref" / "libhilbert_ref.so"


def all_cells():
    g = np.arange(128, dtype=np.uint32)
    i, j, k = np.meshgrid(g, g, g, indexing="ij")
    return i.reshape(-1), j.reshape(-1), k.reshape(-1)


@pytest.fixture(scope="module")
def product_index_fn(tmp_path_factory):
    """Compile the product header into a throw-away host shim (same source the device kernel uses)."""
    d = tmp_path_factory.mktemp("hilbert_shim")
    src = d / "shim.cpp"
    src.write_text(
        '#include "%s"\n'
        'extern "C" void fill(unsigned* out) { for (unsigned i=0;i<128;i++) for (unsigned j=0;j<128;j++) for (unsigned k=0;k<128;k++)'
        " out[(i*128+j)*128+k] = tmb::hilbert3d_index(i,j,k,8); }\n" % (ROOT / "timemachine_b200/csrc/hilbert_curve.h")
    )
    so = d / "shim.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", str(src), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    out = np.empty(128**3, dtype=np.uint32)
    lib.fill(out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)))
    return out


def test_product_header_matches_oracle_exhaustively(product_index_fn):
    i, j, k = all_cells()
    ref = O.hilbert3d_index(i, j, k).astype(np.uint32)
    assert np.array_equal(product_index_fn, ref)
    assert ref.max() == 2**21 - 1  # only 7 of the 8 bits per axis are populated
    assert len(np.unique(ref)) == 128**3


@pytest.mark.skipif(not REF_LIB.exists(), reason="oracle/_ref/libhilbert_ref.so not built (make -C oracle/ref_build)")
def test_matches_reference_vendored_routine_exhaustively(product_index_fn):
    lib = ctypes.CDLL(str(REF_LIB))
    lib.hilbert_c2i.restype = ctypes.c_ulonglong
    lib.hilbert_c2i.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.POINTER(ctypes.c_ulonglong)]
    # spot values quoted in SURVEY.md §8c
    c = (ctypes.c_ulonglong * 3)(0, 0, 43)
    assert lib.hilbert_c2i(3, 8, c) == 33291
    c = (ctypes.c_ulonglong * 3)(127, 127, 127)
    assert lib.hilbert_c2i(3, 8, c) == 1414745
    # exhaustive (2M calls through ctypes would take minutes: sample a deterministic 1-in-13 subset plus all faces)
    rng = np.random.default_rng(0)
    idx = np.unique(np.concatenate([np.arange(0, 128**3, 13), rng.integers(0, 128**3, 20000)]))
    for cell in idx:
        i, j, k = cell // (128 * 128), (cell // 128) % 128, cell % 128
        c = (ctypes.c_ulonglong * 3)(int(i), int(j), int(k))
        assert lib.hilbert_c2i(3, 8, c) == product_index_fn[cell]


def test_oracle_perm_is_stable_and_a_permutation(rng):
    x = rng.uniform(-3, 9, (500, 3))
    box = np.diag([4.0, 5.0, 6.0])
    perm = O.hilbert_perm(x, box)
    assert sorted(perm.tolist()) == list(range(500))
    keys = O.hilbert_keys(x, box)
    assert np.all(np.diff(keys[perm].astype(np.int64)) >= 0)
    # ties keep atom order
    same = np.flatnonzero(np.diff(keys[perm].astype(np.int64)) == 0)
    assert np.all(perm[same] < perm[same + 1])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 33, 1000, 23558])
def test_device_sort_bit_exact(n, rng):
    from timemachine_b200 import custom_ops as ops

    box = np.diag(rng.uniform(3.0, 7.0, 3))
    x = rng.uniform(-2.0, 2.0, (n, 3)) * np.diagonal(box) + rng.normal(0, 5.0, (n, 3))
    perm = ops.HilbertSort(n).sort(x, box)
    assert perm.dtype == np.uint32
    assert np.array_equal(perm, O.hilbert_perm(x, box))


@pytest.mark.gpu
def test_device_sort_compactness(rng):
    """The reference's own test (tests/test_hilbert_sort.py:32-45): sorted blocks are spatially compact."""
    from timemachine_b200 import custom_ops as ops

    n, L = 6000, 4.0
    box = np.eye(3) * L
    x = rng.uniform(0, L, (n, 3))
    perm = ops.HilbertSort(n).sort(x, box)

    def mean_block_extent(c):
        ext = []
        for b in range(0, n - 31, 32):
            blk = c[b : b + 32]
            ext.append(np.linalg.norm(blk.max(0) - blk.min(0)))
        return np.mean(ext)

    assert mean_block_extent(x[perm]) < 0.6 * mean_block_extent(x)
