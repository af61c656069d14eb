"""ctypes access to oracle/liboracle.so (C restatement of batched LU; test infrastructure)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _sfx(dtype):
    return {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dtype)]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _threads():
    return max(1, os.cpu_count() or 1)


def _parallel(fn, batch, threads=None):
    """Run fn(lo, hi) over contiguous slices of the batch on a thread pool."""
    from concurrent.futures import ThreadPoolExecutor

    threads = threads or _threads()
    if batch < 2 * threads or threads == 1:
        fn(0, batch)
        return
    step = (batch + threads - 1) // threads
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda lo: fn(lo, min(batch, lo + step)), range(0, batch, step)))


def lu_factor(A, threads=None):
    A = np.ascontiguousarray(A)
    batch, n = (A.shape[0], A.shape[-1]) if A.ndim == 3 else (1, A.shape[-1])
    lu = np.empty_like(A)
    piv = np.empty(A.shape[:-1], dtype=np.int32)
    f = getattr(lib(), f"oracle_lu_factor_{_sfx(A.dtype)}")
    A3, lu3, piv2 = A.reshape(batch, n, n), lu.reshape(batch, n, n), piv.reshape(batch, n)
    _parallel(lambda lo, hi: f(_p(A3[lo:hi]), _p(lu3[lo:hi]), _p(piv2[lo:hi]), ctypes.c_int64(hi - lo),
                               ctypes.c_int(n)), batch, threads)
    return lu, piv


def lu_solve(lu, piv, b, trans=0, threads=None):
    lu = np.ascontiguousarray(lu)
    b = np.ascontiguousarray(b, dtype=lu.dtype)
    piv = np.ascontiguousarray(piv, dtype=np.int32)
    batch, n = (lu.shape[0], lu.shape[-1]) if lu.ndim == 3 else (1, lu.shape[-1])
    x = np.empty_like(b)
    f = getattr(lib(), f"oracle_lu_solve_{_sfx(lu.dtype)}")
    lu3, piv2, b2, x2 = lu.reshape(batch, n, n), piv.reshape(batch, n), b.reshape(batch, n), x.reshape(batch, n)
    _parallel(lambda lo, hi: f(_p(lu3[lo:hi]), _p(piv2[lo:hi]), _p(b2[lo:hi]), _p(x2[lo:hi]),
                               ctypes.c_int64(hi - lo), ctypes.c_int(n), ctypes.c_int(trans)), batch, threads)
    return x


def lu_factor_solve(A, b, threads=None):
    lu, piv = lu_factor(A, threads)
    return lu_solve(lu, piv, b, 0, threads), lu, piv
