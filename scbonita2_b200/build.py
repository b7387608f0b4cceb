"""Build ``libscbonita_b200.so`` in-tree with nvcc for sm_100a (no torch involved)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(CSRC, "libscbonita_b200.so")
SOURCES = ["scb_api.cu", "scb_pack.cu", "scb_counts.cu", "scb_rules.cu", "scb_sweep.cu", "scb_peer.cu", "scb_dtw.cu"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the scbonita2_b200 CUDA library")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(INCLUDE, "scbonita_b200.h"))
    return any(os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
           "-std=c++17", "-diag-suppress", "128", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static",
           "-I", INCLUDE, "-I", CSRC, "-o", LIB_PATH]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    if os.environ.get("SCB_TRACE"):   # phase timestamps of rule_counts_kernel (tools/trace_counts.py)
        cmd += ["-DSCB_TRACE"]
    if os.environ.get("SCB_DIRECT_RED"):   # experiments
        cmd += ["-DSCB_DIRECT_RED=" + os.environ["SCB_DIRECT_RED"]]
    if os.environ.get("SCB_SWEEP_THREADS"):   # experiments
        cmd += ["-DSCB_SWEEP_THREADS=" + os.environ["SCB_SWEEP_THREADS"]]
    for d in os.environ.get("SCB_DEFINES", "").split():   # experiments: SCB_DEFINES="NAME=VALUE ..."
        cmd += ["-D" + d]
    if os.environ.get("SCB_CONSUMER_WARPS"):   # experiments
        cmd += ["-DSCB_CONSUMER_WARPS=" + os.environ["SCB_CONSUMER_WARPS"]]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
