"""In-tree build of the C-ABI library ``libub200.so`` (sm_100a only, nvcc).

``python -m uncertainty_nerf_gs_b200.build`` or ``build_library()``.  The library is built next to
this file so that it travels with a snapshot of the repository; it is git-ignored.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import List

PKG_DIR = Path(__file__).resolve().parent
CSRC_DIR = PKG_DIR / "csrc"
INCLUDE_DIR = PKG_DIR.parent / "include"
BUILD_DIR = PKG_DIR / "build"
LIB_PATH = PKG_DIR / "libub200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("UB_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the ub200 library can only be built with the CUDA toolkit")
    return exe


def sources() -> List[Path]:
    return sorted(CSRC_DIR.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC_DIR.glob("*.cu")) + list(CSRC_DIR.glob("*.cuh")) + list(INCLUDE_DIR.glob("*.h"))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: Path, log_dir: Path) -> Path:
    obj = BUILD_DIR / (src.stem + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(INCLUDE_DIR), "-c", str(src), "-o", str(obj)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    (log_dir / (src.stem + ".ptxas.log")).write_text(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{proc.stdout}\n{proc.stderr}")
    return obj


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every ``csrc/*.cu`` for sm_100a and link ``libub200.so``.  Skips the work when the
    sources are unchanged since the last successful build."""
    BUILD_DIR.mkdir(exist_ok=True)
    stamp = BUILD_DIR / "stamp.sha256"
    digest = _digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
        objs = list(pool.map(lambda s: _compile_one(s, BUILD_DIR), srcs))
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB_PATH),
           *map(str, objs), "-lcudart"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed:\n{proc.stdout}\n{proc.stderr}")
    stamp.write_text(digest)
    if verbose:
        for s in srcs:
            print((BUILD_DIR / (s.stem + ".ptxas.log")).read_text())
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
