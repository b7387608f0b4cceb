"""Timeline of the GA step's two kernels over a few steps of a CUDA graph (SCB_TRACE build): globaltimer stamps of
CTA 0 -- counting kernel: entry / loop exit / after griddepcontrol.wait; fitness kernel: entry / after wait / end."""
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
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(K):
        plan.fitness_step(planes, lut, out=out, chained=True)
g.replay(); torch.cuda.synchronize()
h = ctypes.CDLL(_lib.LIB_PATH)
g.replay(); torch.cuda.synchronize()
a = np.zeros((64, 4), dtype=np.uint64); b = np.zeros((64, 4), dtype=np.uint64)
h.scb_debug_pdl_a(a.ctypes.data_as(ctypes.c_void_p)); h.scb_debug_pdl_b(b.ctypes.data_as(ctypes.c_void_p))
ev = []
for i in range(64):
    if a[i, 0]: ev += [(int(a[i, 0]), "A entry"), (int(a[i, 1]), "A loop exit"), (int(a[i, 2]), "A after wait")]
    if b[i, 0]: ev += [(int(b[i, 0]), "B entry"), (int(b[i, 1]), "B after wait"), (int(b[i, 2]), "B end")]
ev.sort()
t0 = ev[-40][0]
for t, what in ev[-40:]:
    print(f"{(t - t0) / 1e3:8.2f} us  {what}")
