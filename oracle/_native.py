"""loader of the oracle's C restatement (oracle/csrc/oracle_kernels.c); compiled on first use for the
CPU it runs on (the GPU box's host CPU differs from the build container's).  Test infrastructure."""
from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_lib = None
_tried = False


def cpu_tag() -> str:
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except Exception:
        flags = "generic"
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


def lib():
    global _lib, _tried
    if _lib is not None or _tried:
        return _lib
    _tried = True
    out = _HERE / "_build" / f"liboracle_{cpu_tag()}.so"
    if not out.exists():
        out.parent.mkdir(exist_ok=True)
        try:
            subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared",
                            str(_HERE / "csrc" / "oracle_kernels.c"), "-o", str(out)], check=True, capture_output=True)
        except Exception:
            return None
    try:
        l = C.CDLL(str(out))
    except OSError:
        return None
    l.gemm_i8_i32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]
    l.gemm_i8_i32.restype = None
    l.hash_combine.argtypes = [C.c_uint64, C.c_void_p, C.c_int32]
    l.hash_combine.restype = C.c_uint64
    _lib = l
    return l


def gemm_i8_i32(a8: np.ndarray, w8: np.ndarray):
    l = lib()
    if l is None:
        return None
    a8 = np.ascontiguousarray(a8, dtype=np.int8)
    w8 = np.ascontiguousarray(w8, dtype=np.int8)
    M, K = a8.shape
    N = w8.shape[0]
    out = np.empty((M, N), dtype=np.int32)
    l.gemm_i8_i32(a8.ctypes.data, w8.ctypes.data, out.ctypes.data, M, N, K)
    return out


def hash_combine(prev: int, vec) -> int | None:
    l = lib()
    if l is None:
        return None
    v = np.ascontiguousarray(vec, dtype=np.int32)
    return int(l.hash_combine(prev, v.ctypes.data, len(v)))
