"""Time scb_ko_ki_sweep on the H1M shape (for A/B experiments)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tools import synth
from scbonita2_b200 import device
from scbonita2_b200.importance_scores import compile_network

cells = int(os.environ.get("CELLS", 1_000_000))
specs = synth.random_network(200, 0)
planes_np = synth.relax_and_flip(specs, synth.random_planes(200, cells, 1000), cells, 0, relax_steps=30, flip=0.05)
nodes = bench.make_nodes(specs)
planes = device.BitPlanes.from_packed(planes_np, cells)
par, lut = compile_network(nodes)
iters = int(os.environ.get("ITERS", 3))
for i in range(iters):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = device.ko_ki_sweep(planes, par, lut, stats=(i == iters - 1))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"ring={os.environ.get('SCB_SWEEP_RING','auto')} sweep {dt*1e3:.1f} ms  raw {int(out[0].sum().item())}  {device.LAST_SWEEP_STATS}")
