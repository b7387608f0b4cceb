"""Phase times of population_fitness_kernel (CTA 0, SCB_TRACE build) inside a graph of GA steps."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import synth
from scbonita2_b200 import device, _lib

n, cells, pop, K = 200, 1_000_000, 4096, 12
specs = synth.random_network(n, 0)
planes = device.BitPlanes.from_packed(synth.random_planes(n, cells, 1000), cells)
plan = device.RuleCountsPlan(n, [s.index for s in specs], [synth.spec_parent_rows(s) for s in specs])
lut = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (pop, n), dtype=np.uint8)).cuda()
out = torch.empty(pop, dtype=torch.int64, device="cuda")
for _ in range(3):
    plan.fitness_step(planes, lut, out=out, chained=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(K):
        plan.fitness_step(planes, lut, out=out, chained=True)
g.replay(); g.replay(); torch.cuda.synchronize()
h = ctypes.CDLL(_lib.LIB_PATH)
t = np.zeros(16, dtype=np.uint64)
h.scb_debug_fit_phases(t.ctypes.data_as(ctypes.c_void_p))
names = ["entry->wait", "wait->sums in smem", "sums->moebius done", "->base/wide", "->nibble tables", "->population loop end"]
for i, name in enumerate(names):
    print(f"{name:26s} {(int(t[i + 1]) - int(t[i])) / 1e3:6.2f} us")
print(f"inside the inversion: sums gathered at +{(int(t[8]) - int(t[2])) / 1e3:.2f} us, written at +{(int(t[9]) - int(t[2])) / 1e3:.2f} us, barrier passed at +{(int(t[3]) - int(t[2])) / 1e3:.2f} us")
