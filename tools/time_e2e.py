"""Phase times of the host-buffer GA step (scb_host_load_planes / scb_host_rule_counts / scb_host_population_fitness)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import synth
from scbonita2_b200 import device

n, cells, pop = 200, 1_000_000, 4096
specs = synth.random_network(n, 0)
planes_np = synth.random_planes(n, cells, 1000)
n_words = (cells + 31) // 32
packed = torch.from_numpy(np.ascontiguousarray(planes_np[:, :n_words]).view(np.int32)).pin_memory().numpy().view(np.uint32)
lut = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (pop, n), dtype=np.uint8)).pin_memory().numpy()
targets = np.array([s.index for s in specs], dtype=np.int32)
parents = np.array([synth.spec_parent_rows(s) for s in specs], dtype=np.int32)
ws = device.HostWorkspace()
ph = np.zeros(3)
for it in range(13):
    t0 = time.perf_counter(); ws.load_planes(packed, cells)
    t1 = time.perf_counter(); ws.rule_counts(targets, parents, want=False)
    t2 = time.perf_counter(); out, _ = ws.population_fitness(lut)
    t3 = time.perf_counter()
    if it >= 3: ph += [t1 - t0, t2 - t1, t3 - t2]
print("load_planes %.1f us, rule_counts %.1f us, population_fitness %.1f us (mean of 10)" % tuple(ph / 10 * 1e6), int(out.sum()))
