"""Time scb_rule_counts alone (CUDA-graph replay, L2 flushed between replays, no host sync inside
the timed region) for quick A/B experiments."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import synth
from scbonita2_b200 import device

def main():
    n = int(os.environ.get("N_NODES", 200)); cells = int(os.environ.get("N_CELLS", 1_000_000))
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
    for cold in (True, False):
        pairs = []
        torch.cuda.synchronize()
        for _ in range(30):
            if cold:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            pairs.append((e0, e1))
        torch.cuda.synchronize()
        ts = [a.elapsed_time(b) * 1e3 for a, b in pairs]
        print(f"stages={os.environ.get('SCB_COUNTS_STAGES','max')} {'cold' if cold else 'warm'}: counts graph replay us: "
              f"median {np.median(ts):.1f} mean {np.mean(ts):.1f} min {min(ts):.1f}  checksum {int(plan.counts.sum().item())}")

main()
