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
_BS_PATH = os.path.join(_HERE, "_build", "libscb_oracle_bs.so")
_LIB = None
_BS = None


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


def lib_bs():
    """The 64-cells-per-word restatement (``oracle/c/scb_oracle_bs.c``), for 10^5..10^6 cells."""
    global _BS
    if _BS is None:
        src = os.path.join(_HERE, "c", "scb_oracle_bs.c")
        if not os.path.exists(_BS_PATH) or os.path.getmtime(_BS_PATH) < os.path.getmtime(src):
            subprocess.run(["make", "-C", os.path.join(_HERE, "c"), "-s", _BS_PATH.replace(_HERE, "..")], check=True)
        L = ctypes.CDLL(_BS_PATH)
        L.oracle_bs_importance_raw.restype = ctypes.c_int64
        _BS = L
    return _BS


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


def planes64(planes32: np.ndarray, n_cells: int) -> np.ndarray:
    """uint64 [n][ceil(n_cells/64)] from uint32 cell bitplanes (bit b of word w = cell 32w+b)."""
    planes32 = np.ascontiguousarray(planes32, dtype=np.uint32)
    n, have = planes32.shape
    w32 = (n_cells + 31) // 32
    w64 = (n_cells + 63) // 64
    buf = np.zeros((n, 2 * w64), dtype=np.uint32)
    buf[:, :w32] = planes32[:, :w32]
    return buf.view(np.uint64)        # little-endian host: word pair (lo, hi) -> one uint64


def importance_raw_bitsliced(planes32: np.ndarray, n_cells: int, parents, luts, steps: int = 50, threads: int = 0,
                             with_stats: bool = False):
    """(raw[N], missing[2N+1], keyerror) from the 64-cells-per-word restatement; ``planes32`` are the
    same uint32 bitplanes the CUDA path takes.  ``threads`` = 0: every host core."""
    p64 = planes64(planes32, n_cells)
    n = p64.shape[0]
    par = np.ascontiguousarray(parents, dtype=np.int32).reshape(n, 3)
    lut = np.ascontiguousarray(luts, dtype=np.uint8).reshape(n)
    raw = np.zeros(n, dtype=np.int64)
    missing = np.zeros(2 * n + 1, dtype=np.int64)
    stats = np.zeros(2, dtype=np.int64)
    key = lib_bs().oracle_bs_importance_raw(
        _p(p64, ctypes.c_uint64), ctypes.c_int(n), ctypes.c_int64(p64.shape[1]), ctypes.c_int64(n_cells),
        _p(par, ctypes.c_int32), _p(lut, ctypes.c_uint8), ctypes.c_int(steps),
        ctypes.c_int(threads or os.cpu_count() or 1), _p(raw, ctypes.c_int64), _p(missing, ctypes.c_int64),
        _p(stats, ctypes.c_int64))
    if with_stats:
        return raw, missing, int(key), (int(stats[0]), int(stats[1]))
    return raw, missing, int(key)
