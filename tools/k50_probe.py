"""Per-pathway sweep times and pass statistics on K50-shaped inputs (diagnostic)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tools import synth
from scbonita2_b200 import device
from scbonita2_b200.importance_scores import compile_network

cells = int(os.environ.get("CELLS", 500_000))
rng = np.random.default_rng(0)
sizes = rng.integers(30, 201, 50)
for p in [int(x) for x in os.environ.get("PATHWAYS", "0,1,2,3,4,5").split(",")]:
    n = int(sizes[p])
    specs = synth.random_network(n, p)
    start = synth.random_planes(n, cells, 1000 + p)
    planes_np = synth.relax_and_flip(specs, start, cells, p, relax_steps=int(os.environ.get("RELAX", 10)), flip=0.05)
    nodes = bench.make_nodes(specs)
    planes = device.BitPlanes.from_packed(planes_np, cells)
    par, lut = compile_network(nodes)
    for i in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = device.ko_ki_sweep(planes, par, lut, stats=(i == 1))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    s = device.LAST_SWEEP_STATS
    print(f"pathway {p}: N={n} sweep {dt*1e3:.1f} ms  T={s['fast_steps']} deep={s['deep_ring']} rechecked {s['rechecked_cells']/s['cell_node_pairs']:.3f} "
          f"cleanup {s['deferred_cells']/s['cell_node_pairs']:.3f} missing_normal {int(out[1][0].item())}")
