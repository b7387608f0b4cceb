"""Eager (no graph) timing of scb_rule_counts_run: 20 back-to-back launches per measurement; for
debug-mode experiments with an SCB_TRACE build (SCB_COUNTS_DBG=1 no arithmetic, =2 no copies)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import synth
from scbonita2_b200 import device

n = int(os.environ.get("N_NODES", 200)); cells = int(os.environ.get("N_CELLS", 4_000_000))
specs = synth.random_network(n, 0)
planes = device.BitPlanes.from_packed(synth.random_planes(n, cells, 1000), cells)
plan = device.RuleCountsPlan(n, [s.index for s in specs], [synth.spec_parent_rows(s) for s in specs])
for _ in range(3):
    plan.run(planes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    plan.run(planes)
e1.record(); torch.cuda.synchronize()
print(f"dbg={os.environ.get('SCB_COUNTS_DBG','0')} inflight={os.environ.get('SCB_COUNTS_INFLIGHT','-')} cells={cells}: {e0.elapsed_time(e1)*1e3/20:.1f} us per launch (20 back to back, warm)")
