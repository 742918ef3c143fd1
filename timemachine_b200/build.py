"""Build libtmb200.so (all CUDA kernels + host classes + C ABI) for sm_100a with nvcc, in-tree.

    python -m timemachine_b200.build            # incremental
    python -m timemachine_b200.build --force    # rebuild everything

nvcc cross-compiles without a GPU.  Objects go to timemachine_b200/csrc/build/, the library to
timemachine_b200/lib/libtmb200.so (git-ignored, shipped to the GPU box by gpurun).
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
SRC_DIR = PKG_DIR / "csrc"
OBJ_DIR = SRC_DIR / "build"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libtmb200.so"

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

# --fmad=false: every fused multiply-add in the kernels is written explicitly (see csrc/nb_math.cuh); the compiler
# never contracts on its own, so a pair term is rounded identically in every kernel that evaluates it.
NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--fmad=false",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-Wall",
    "-Xcudafe",
    "--diag_suppress=177",  # unused-variable noise from templated kernels
]


def _sources() -> list[Path]:
    return sorted(SRC_DIR.glob("*.cu"))


def _headers_digest() -> str:
    import hashlib

    hs = sorted(list(SRC_DIR.glob("*.cuh")) + list(SRC_DIR.glob("*.hpp")) + list(SRC_DIR.glob("*.h")))
    hs.append(PKG_DIR.parent / "include" / "tmb200.h")
    h = hashlib.sha256()
    for f in hs:
        if f.exists():
            h.update(f.name.encode())
            h.update(f.read_bytes())
    return h.hexdigest()


def _stamp(src: Path, headers: str, flags: list[str]) -> str:
    """What an object file was built from: source text, every header, the flags.  Content, not mtimes: the snapshot that
    travels to the GPU box does not keep the relative order of file times, and a needless rebuild there costs a minute."""
    import hashlib

    h = hashlib.sha256()
    h.update(src.read_bytes())
    h.update(headers.encode())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


# barostat.cu and exchange.cu inline cbrtf / logf / expf from libdevice and must contract them like the reference build
# does (default fmad); all of their own arithmetic is written with explicit round-to-nearest intrinsics.
DEFAULT_FMAD_SOURCES = {"barostat.cu", "exchange.cu"}


def _compile(src: Path, extra: list[str]) -> tuple[Path, str]:
    obj = OBJ_DIR / (src.stem + ".o")
    flags = [f for f in NVCC_FLAGS if not (src.name in DEFAULT_FMAD_SOURCES and f == "--fmad=false")]
    cmd = [NVCC, *flags, *extra, "-c", str(src), "-o", str(obj)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{proc.stdout}\n{proc.stderr}")
    return obj, proc.stderr


def build_library(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> Path:
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    srcs = _sources()
    headers = _headers_digest()
    extra = ["-Xptxas", "-v"] if ptxas_info else []
    extra += os.environ.get("TMB_NVCC_EXTRA", "").split()  # experiments, e.g. -DCQ_MIN_CTAS=3
    todo = []
    stamps = {}
    for s in srcs:
        obj = OBJ_DIR / (s.stem + ".o")
        stamp_file = OBJ_DIR / (s.stem + ".stamp")
        flags = [f for f in NVCC_FLAGS if not (s.name in DEFAULT_FMAD_SOURCES and f == "--fmad=false")] + extra
        stamps[s] = _stamp(s, headers, flags)
        if force or ptxas_info or not obj.exists() or not stamp_file.exists() or stamp_file.read_text() != stamps[s]:
            todo.append(s)
    # the library carries a stamp of everything it was built from: where only the .so travelled (the GPU box gets no object
    # files) an up-to-date library is recognised without recompiling
    import hashlib

    lib_stamp = hashlib.sha256("".join(stamps[s] for s in srcs).encode()).hexdigest()
    lib_stamp_file = LIB_DIR / "libtmb200.stamp"
    if not force and not ptxas_info and LIB_PATH.exists() and lib_stamp_file.exists() and lib_stamp_file.read_text() == lib_stamp:
        return LIB_PATH
    if todo:
        if verbose:
            print(f"[tmb200] compiling {len(todo)} translation units with {NVCC}", file=sys.stderr)
        errors = []

        def job(s):
            try:
                return _compile(s, extra)
            except RuntimeError as e:  # keep going so one run reports every broken translation unit
                errors.append(str(e))
                return None

        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            for s, res in zip(todo, pool.map(job, todo)):
                if res:
                    (OBJ_DIR / (s.stem + ".stamp")).write_text(stamps[s])
                if res and (verbose or ptxas_info) and res[1].strip():
                    print(res[1], file=sys.stderr)
        if errors:
            raise RuntimeError("\n".join(errors))
    objs = [OBJ_DIR / (s.stem + ".o") for s in srcs]
    if todo or not LIB_PATH.exists():
        # cuRAND: the barostat draws the reference's XORWOW stream (same seed -> same accept/reject decisions)
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *map(str, objs), "-lcurand"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"link failed:\n{proc.stdout}\n{proc.stderr}")
    lib_stamp_file.write_text(lib_stamp)
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    print(path)
