"""ctypes wrapper around oracle/spmm_ref.c.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_spmm.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "spmm_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        # -march=native is avoided: the .so travels to a different host (GPU box)
        subprocess.check_call(["gcc", "-O3", "-mavx2", "-mfma", "-fopenmp", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_spmm_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def spmm_f32(rowptr: np.ndarray, col: np.ndarray, X: np.ndarray, vals=None) -> np.ndarray:
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, d = rowptr.shape[0] - 1, X.shape[1]
    Y = np.empty((n, d), dtype=np.float32)
    v = None if vals is None else _p(np.ascontiguousarray(vals, dtype=np.float32))
    lib().oracle_spmm_csr_f32(_p(rowptr), _p(col), v, _p(X), ctypes.c_int64(d), _p(Y), ctypes.c_int64(d),
                              ctypes.c_int64(n), ctypes.c_int32(d))
    return Y


def spmm_f64acc(rowptr: np.ndarray, col: np.ndarray, X: np.ndarray) -> np.ndarray:
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, d = rowptr.shape[0] - 1, X.shape[1]
    Y = np.empty((n, d), dtype=np.float64)
    lib().oracle_spmm_csr_f64acc(_p(rowptr), _p(col), _p(X), ctypes.c_int64(d), _p(Y), ctypes.c_int64(d),
                                 ctypes.c_int64(n), ctypes.c_int32(d))
    return Y


def threads() -> int:
    return int(lib().oracle_spmm_threads())
