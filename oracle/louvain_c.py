"""ctypes wrapper for oracle/louvain_ref.c (TEST INFRASTRUCTURE -- see oracle/__init__.py)."""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liblouvain_ref.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "louvain_ref.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            build()
        _lib = ctypes.CDLL(_LIB)
        _lib.louvain_ref.restype = ctypes.c_int64
        _lib.louvain_ref.argtypes = [
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_double, ctypes.c_uint64, ctypes.c_void_p,
        ]
        _lib.louvain_ref_parallel0.restype = ctypes.c_int64
        _lib.louvain_ref_parallel0.argtypes = [
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_uint64, ctypes.c_void_p,
        ]
        _lib.louvain_ref_parallel0_w.restype = ctypes.c_int64
        _lib.louvain_ref_parallel0_w.argtypes = [
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_uint64, ctypes.c_void_p,
        ]
    return _lib


def louvain(indptr, indices, weights=None, resolution=1.0, seed=0, level0="sequential"):
    """Same contract as oracle.louvain_ref.louvain."""
    lib = _load()
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    n = indptr.size - 1
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    out = np.empty(max(n, 1), dtype=np.int64)
    if level0 == "parallel":
        if w is not None:  # fixed-point first level (oracle/louvain_ref.py:level0_parallel with weights)
            lib.louvain_ref_parallel0_w(n, indptr.ctypes.data, indices.ctypes.data, w.ctypes.data, float(resolution),
                                        int(seed) & ((1 << 64) - 1), out.ctypes.data)
        else:
            lib.louvain_ref_parallel0(n, indptr.ctypes.data, indices.ctypes.data, float(resolution),
                                      int(seed) & ((1 << 64) - 1), out.ctypes.data)
        return out[:n]
    lib.louvain_ref(
        n, indptr.ctypes.data, indices.ctypes.data, None if w is None else w.ctypes.data,
        float(resolution), int(seed) & ((1 << 64) - 1), out.ctypes.data,
    )
    return out[:n]
