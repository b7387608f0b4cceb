"""GA step timing: the fused pair (sums kernel + fitness) and the round-1 pair (counts + fitness), each
as one CUDA graph of K steps over rotating >L2 inputs, plus each counting kernel alone."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import synth
from scbonita2_b200 import device

n = int(os.environ.get("N_NODES", 200)); cells = int(os.environ.get("N_CELLS", 1_000_000)); K = 20
specs = synth.random_network(n, 0)
planes = device.BitPlanes.from_packed(synth.random_planes(n, cells, 1000), cells)
n_rot = max(2, -(-2 * 126 * 1024 * 1024 // (planes.words.numel() * 4)) + 1) if cells <= 4_000_000 else 2
rot = [planes] + [device.BitPlanes(planes.words.clone(), planes.n_rows, planes.n_cells) for _ in range(n_rot - 1)]
plan = device.RuleCountsPlan(n, [s.index for s in specs], [synth.spec_parent_rows(s) for s in specs])
lut = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (4096, n), dtype=np.uint8)).cuda()
out = torch.empty(4096, dtype=torch.int64, device="cuda")

def graph_of(body):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        body()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    return g

def timed(g):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K

fused = graph_of(lambda: [plan.fitness_step(rot[k % n_rot], lut, out=out, chained=True) for k in range(K)])
t_fused = timed(fused); ref = out.clone()
old = graph_of(lambda: [device.population_fitness(plan.run(rot[k % n_rot], chained=True), lut, out=out) for k in range(K)])
t_old = timed(old); assert os.environ.get("SCB_SUMS_DBG", "0") != "0" or torch.equal(ref, out)
sums = graph_of(lambda: [plan.sums_run(rot[k % n_rot], chained=True) for k in range(K)])
t_sums = timed(sums); plan.scratch.zero_()
cnts = graph_of(lambda: [plan.run(rot[k % n_rot], chained=True) for k in range(K)])
t_cnts = timed(cnts)
print(f"cells={cells} nodes={n}: fused step {t_fused:.1f} us, round-1 step {t_old:.1f} us; sums kernel alone {t_sums:.1f} us, counts kernel alone {t_cnts:.1f} us")
plan._sums_cells = cells
fs = graph_of(lambda: [plan.fitness_from_sums(lut, out=out) for k in range(K)])
t_fs = timed(fs)
counts = plan.run(rot[0]).clone()
fc = graph_of(lambda: [device.population_fitness(counts, lut, out=out) for k in range(K)])
t_fc = timed(fc)
print(f"fitness kernel alone: from sums {t_fs:.1f} us, from counts {t_fc:.1f} us")
