"""Phase totals of sweep_kernel per mode (needs a library built with SCB_TRACE=1):
SCB_TRACE=1 python -m scbonita2_b200.build --force && python tools/trace_sweep.py"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tools import synth
from scbonita2_b200 import device, _lib
from scbonita2_b200.importance_scores import compile_network

cells = int(os.environ.get("CELLS", 1_000_000))
specs = synth.random_network(200, 0)
planes_np = synth.relax_and_flip(specs, synth.random_planes(200, cells, 1000), cells, 0, relax_steps=30, flip=0.05)
nodes = bench.make_nodes(specs)
planes = device.BitPlanes.from_packed(planes_np, cells)
par, lut = compile_network(nodes)
h = ctypes.CDLL(_lib.lib()._name)
buf = np.zeros((4, 16), dtype=np.uint64)
device.ko_ki_sweep(planes, par, lut)
h.scb_debug_sweep_phases(buf.ctypes.data_as(ctypes.c_void_p), 1)
device.ko_ki_sweep(planes, par, lut)
h.scb_debug_sweep_phases(buf.ctypes.data_as(ctypes.c_void_p), 1)
names = ["setup", "load", "fast ring", "defer", "slow path", "score/output"]
for mode, mname in enumerate(["normal", "koki", "cleanup", "recheck"]):
    n = int(buf[mode, 15])
    if not n:
        continue
    print(f"{mname}: {n} CTAs/batches; mean cycles per phase:",
          ", ".join(f"{names[i]} {buf[mode, i] / n:.0f}" for i in range(6)))
print("cleanup cells by outcome: no repeat %d, lambda<=2 %d, 3-4 %d, 5-8 %d, >8 %d" % tuple(int(x) for x in buf[2, 8:13]))
for mode, mname in ((1, "koki"), (3, "recheck"), (2, "cleanup")):
    print(f"{mname}: scored lanes {int(buf[mode, 13])}, of them cycling (either run) {int(buf[mode, 14])}")
