"""ctypes loader of the C restatement (``oracle/c/scb_oracle.c``) — TEST INFRASTRUCTURE ONLY.

The C port follows the same reference routines as ``ref_port`` (it is checked against it in
``tests/test_oracle_c.py``) and exists so that mid-size parity cases finish in seconds."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libscb_oracle.so")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        src = os.path.join(_HERE, "c", "scb_oracle.c")
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
            subprocess.run(["make", "-C", os.path.join(_HERE, "c"), "-s"], check=True)
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_rule_difference.restype = ctypes.c_int64
        L.oracle_importance_raw.restype = ctypes.c_int64
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def rule_difference(dense01: np.ndarray, target: int, parents, lut: int) -> int:
    dense01 = np.ascontiguousarray(dense01, dtype=np.uint8)
    par = np.ascontiguousarray(parents, dtype=np.int32)
    return int(lib().oracle_rule_difference(_p(dense01, ctypes.c_uint8), ctypes.c_int64(dense01.shape[1]),
                                            ctypes.c_int(int(target)), _p(par, ctypes.c_int32),
                                            ctypes.c_uint8(int(lut))))


def importance_raw(dense01: np.ndarray, parents, luts, steps: int = 50):
    """(raw[N], missing[2N+1], keyerror) — same meaning as ``scb_ko_ki_sweep``'s outputs."""
    dense01 = np.ascontiguousarray(dense01, dtype=np.uint8)
    n = dense01.shape[0]
    par = np.ascontiguousarray(parents, dtype=np.int32).reshape(n, 3)
    lut = np.ascontiguousarray(luts, dtype=np.uint8).reshape(n)
    raw = np.zeros(n, dtype=np.int64)
    missing = np.zeros(2 * n + 1, dtype=np.int64)
    key = lib().oracle_importance_raw(_p(dense01, ctypes.c_uint8), ctypes.c_int(n),
                                      ctypes.c_int64(dense01.shape[1]), _p(par, ctypes.c_int32),
                                      _p(lut, ctypes.c_uint8), ctypes.c_int(steps),
                                      _p(raw, ctypes.c_int64), _p(missing, ctypes.c_int64))
    return raw, missing, int(key)


def trajectory(column01, parents, luts, steps: int = 20) -> np.ndarray:
    start = np.ascontiguousarray(column01, dtype=np.uint8).reshape(-1)
    n = start.size
    par = np.ascontiguousarray(parents, dtype=np.int32).reshape(n, 3)
    lut = np.ascontiguousarray(luts, dtype=np.uint8).reshape(n)
    out = np.zeros((steps, n), dtype=np.uint8)
    lib().oracle_trajectory(_p(start, ctypes.c_uint8), ctypes.c_int(n), _p(par, ctypes.c_int32),
                            _p(lut, ctypes.c_uint8), ctypes.c_int(steps), _p(out, ctypes.c_uint8))
    return out
