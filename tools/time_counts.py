"""Time scb_rule_counts alone (CUDA-graph replay, L2 flushed) for quick A/B experiments."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import synth
from scbonita2_b200 import device

def main():
    n, cells = 200, 1_000_000
    specs = synth.random_network(n, 0)
    planes_np = synth.random_planes(n, cells, 1000)
    planes = device.BitPlanes.from_packed(planes_np, cells)
    targets = [s.index for s in specs]
    parents = [synth.spec_parent_rows(s) for s in specs]
    plan = device.RuleCountsPlan(n, targets, parents)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        plan.run(planes)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        plan.run(planes)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        plan.run(planes)
    ts = []
    for _ in range(20):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"ctas_per_sm={os.environ.get('SCB_COUNTS_CTAS_PER_SM','max')}: counts graph replay us: median {np.median(ts):.1f} min {min(ts):.1f}  checksum {int(plan.counts.sum().item())}")

main()
