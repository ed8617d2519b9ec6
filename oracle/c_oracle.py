"""ctypes loader of oracle/_build/liboracle_flat_ip.so — TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle_flat_ip.so"
_lib = None


def build(force: bool = False) -> Path:
    if force or not _SO.exists() or _SO.stat().st_mtime < (_HERE / "flat_ip.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        _lib = ctypes.CDLL(str(_SO))
        _lib.oracle_flat_ip_search.restype = ctypes.c_int
        _lib.oracle_flat_ip_search.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_int]
        _lib.oracle_normalize_rows.restype = None
        _lib.oracle_normalize_rows.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    return _lib


def search(xn: np.ndarray, qn: np.ndarray, k: int, nthreads: int = 0):
    xn = np.ascontiguousarray(xn, np.float32)
    qn = np.ascontiguousarray(qn, np.float32)
    n, d = xn.shape
    nq = qn.shape[0]
    k = min(k, n)
    s = np.empty((nq, k), np.float32)
    i = np.empty((nq, k), np.int64)
    _load().oracle_flat_ip_search(xn.ctypes.data, n, d, qn.ctypes.data, nq, k, s.ctypes.data, i.ctypes.data,
                                  nthreads or (os.cpu_count() or 1))
    return s, i


def normalize_rows(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    _load().oracle_normalize_rows(x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data)
    return out
