"""Builds csrc/*.cu into lib/libtt_b200.so for sm_100a (in-tree, so the .so travels with the repo).

    python two-tower-model-v2_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The library links the CUDA runtime statically and resolves
cuTensorMapEncodeTiled at run time through cudaGetDriverEntryPoint, so it loads (dlopen) on a
box without a driver; every entry point then returns TT_ERR_CUDA.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "lib" / "libtt_b200.so"
INCLUDE = HERE.parent / "include"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _deps():
    return sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))


def _stale(target: Path, inputs) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(p.stat().st_mtime > t for p in inputs)


def _compile(src: Path, force: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    if force or _stale(obj, [src] + _deps()):
        cmd = [NVCC, *ARCH, *CFLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return obj


def build_library(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), srcs))
    if force or _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
