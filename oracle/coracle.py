"""ctypes binding to oracle/_build/liboracle.so (TEST INFRASTRUCTURE; see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build() -> str:
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        _lib.oracle_pairwise_distance.argtypes = [p, p, p, i64, i64, i64, ctypes.c_int,
                                                  ctypes.c_double, i64, i64]
        _lib.oracle_fused_l2_nn.argtypes = [p, p, p, p, i64, i64, i64, ctypes.c_int, i64, i64]
        _lib.oracle_l2_expanded_f32.argtypes = [p, p, p, i64, i64, i64, ctypes.c_int, i64, i64]
    return _lib


def _run_rows(fn, m, threads):
    threads = max(1, min(int(threads), m))
    if threads == 1:
        fn(0, m)
        return
    cuts = [m * t // threads for t in range(threads + 1)]
    ts = [threading.Thread(target=fn, args=(cuts[t], cuts[t + 1])) for t in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def pairwise_distance(x, y, metric, metric_arg=2.0, threads=1):
    x, y = _f32(x), _f32(y)
    m, k = x.shape
    n = y.shape[0]
    out = np.empty((m, n), dtype=np.float64)
    L = lib()
    _run_rows(lambda a, b: L.oracle_pairwise_distance(x.ctypes.data, y.ctypes.data, out.ctypes.data,
                                                      m, n, k, int(metric), float(metric_arg), a, b),
              m, threads)
    return out


def fused_l2_nn(x, y, sqrt=False, threads=1):
    x, y = _f32(x), _f32(y)
    m, k = x.shape
    n = y.shape[0]
    idx = np.empty(m, dtype=np.int32)
    val = np.empty(m, dtype=np.float64)
    L = lib()
    _run_rows(lambda a, b: L.oracle_fused_l2_nn(x.ctypes.data, y.ctypes.data, idx.ctypes.data,
                                                val.ctypes.data, m, n, k, int(bool(sqrt)), a, b),
              m, threads)
    return idx, val


def l2_expanded_f32(x, y, sqrt=False, threads=1, out=None):
    x, y = _f32(x), _f32(y)
    m, k = x.shape
    n = y.shape[0]
    if out is None:
        out = np.empty((m, n), dtype=np.float32)
    L = lib()
    _run_rows(lambda a, b: L.oracle_l2_expanded_f32(x.ctypes.data, y.ctypes.data, out.ctypes.data,
                                                    m, n, k, int(bool(sqrt)), a, b), m, threads)
    return out
