"""ctypes binding of ``libscbonita_b200.so`` (the C-ABI declared in ``include/scbonita_b200.h``).

There is no CPU fallback: if the shared library is missing, :func:`lib` raises, and every
product entry point that needs the GPU goes through :func:`lib`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int32, c_int64, c_uint8, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libscbonita_b200.so")

_I32P = POINTER(c_int32)
_I64P = POINTER(c_int64)
_U8P = POINTER(c_uint8)
_U32P = POINTER(c_uint32)

SCB_COUNTS_SCRATCH_CLEAN = 1
SCB_COUNTS_CHAINED = 2

# name -> (restype, argtypes).  Device pointers are passed as integers (c_void_p).
SIGNATURES = {
    "scb_abi_version": (c_int, []),
    "scb_last_error": (c_char_p, []),
    "scb_device_sm_count": (c_int, []),
    "scb_pack_bitplanes": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "scb_binarize_pack_f32": (c_int, [c_void_p, c_int, c_int64, c_int64, ctypes.c_float, c_void_p, c_int64, c_void_p]),
    "scb_binarize_pack_f64": (c_int, [c_void_p, c_int, c_int64, c_int64, ctypes.c_double, c_void_p, c_int64, c_void_p]),
    "scb_unpack_bitplanes": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "scb_chunk_majority": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int64, c_int, c_int,
                                   c_void_p, c_int64, c_void_p]),
    "scb_rule_counts_scratch_bytes": (c_int64, [c_int, c_int]),
    "scb_rule_counts": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                c_void_p, c_int64, c_void_p, c_void_p]),
    "scb_rule_counts_ex": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                   c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    "scb_rule_counts_plan_bytes": (c_int64, [c_int, c_int]),
    "scb_rule_counts_plan_build": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "scb_rule_counts_run": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int64,
                                    c_void_p, c_int, c_void_p]),
    "scb_rule_sums_supported": (c_int, [c_int, c_int, c_int]),
    "scb_rule_sums_run": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_int64, c_int,
                                  c_void_p, c_int, c_int, c_void_p]),
    "scb_population_fitness_sums": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_int, c_int, c_int64,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "scb_rule_error_table": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "scb_rule_error_table_direct": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p]),
    "scb_population_fitness": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p]),
    "scb_peer_region_bytes": (c_int64, [c_int, c_int64]),
    "scb_peer_region_alloc": (c_int, [c_int64, POINTER(c_void_p), c_char_p]),
    "scb_peer_region_open": (c_int, [c_char_p, POINTER(c_void_p)]),
    "scb_peer_region_close": (c_int, [c_void_p]),
    "scb_peer_region_free": (c_int, [c_void_p]),
    "scb_peer_push_counts": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "scb_rule_counts_run_push": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int64,
                                         c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "scb_population_fitness_peers": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p]),
    "scb_peer_error": (c_int, [c_void_p, POINTER(c_int)]),
    "scb_peer_set_timeout": (c_int, [c_void_p, c_uint32]),
    "scb_bitstring_fitness": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "scb_ko_ki_sweep_scratch_bytes": (c_int64, [c_int, c_int64, c_int]),
    "scb_ko_ki_sweep": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int,
                                c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "scb_trajectories": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                 c_void_p, c_int, c_void_p, c_void_p]),
    "scb_trajectories_forced": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                        c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "scb_dtw_table_build": (c_int, [c_int, c_int, c_int, c_void_p]),
    "scb_trajectory_patterns": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "scb_pattern_pair_distances": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "scb_workspace_create": (c_int, [POINTER(c_void_p)]),
    "scb_workspace_destroy": (None, [c_void_p]),
    "scb_host_load_dense": (c_int, [c_void_p, _U8P, c_int, c_int64]),
    "scb_host_load_planes": (c_int, [c_void_p, _U32P, c_int, c_int64]),
    "scb_host_rule_counts": (c_int, [c_void_p, c_int, _I32P, _I32P, _I64P]),
    "scb_host_rule_error_table": (c_int, [c_void_p, c_int64, _I32P, _U8P, _I64P]),
    "scb_host_population_fitness": (c_int, [c_void_p, c_int64, _U8P, _U8P, _I64P, _I64P]),
    "scb_host_ko_ki_sweep": (c_int, [c_void_p, _I32P, _U8P, c_int, _I64P, _I64P, _I64P]),
}

_LIB = None


class ScbError(RuntimeError):
    """A C-ABI call returned a negative SCB_ERR_* code."""


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m scbonita2_b200.build` "
                "(scbonita2_b200 has no CPU fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.scb_abi_version() != 1:
            raise RuntimeError("libscbonita_b200.so ABI version mismatch; rebuild it")
        _LIB = handle
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().scb_last_error().decode("utf-8", "replace")
        raise ScbError(f"{what or 'scbonita_b200'} failed (code {rc}): {msg}")
