"""Small driver for ncu captures: runs each hot kernel a few times on an H1M-shaped input.

    python tools/profile_kernels.py [--cells 1000000] [--nodes 200] [--reps 3] [--what counts,fitness,sweep]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tools import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=1_000_000)
    ap.add_argument("--nodes", type=int, default=200)
    ap.add_argument("--population", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--what", default="counts,fitness,sweep")
    ap.add_argument("--relax", type=int, default=30)
    args = ap.parse_args()
    import torch

    import bench
    from scbonita2_b200 import device, rule_compiler
    from scbonita2_b200.importance_scores import compile_network

    specs = synth.random_network(args.nodes, 0)
    start = synth.random_planes(args.nodes, args.cells, 1000)
    planes_np = synth.relax_and_flip(specs, start, args.cells, 0, relax_steps=args.relax, flip=0.05)
    nodes = bench.make_nodes(specs)
    planes = device.BitPlanes.from_packed(planes_np, args.cells)
    targets = np.array([n.index for n in nodes], dtype=np.int32)
    parents = np.array([rule_compiler.parent_rows(n) for n in nodes], dtype=np.int32)
    what = set(args.what.split(","))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    if "counts" in what or "fitness" in what:
        plan = device.RuleCountsPlan(args.nodes, targets, parents)
        lut = torch.from_numpy(bench.make_population(nodes, args.population, 0)).cuda()
        out = torch.empty(args.population, dtype=torch.int64, device="cuda")
        for _ in range(args.reps):
            flush.fill_(1)
            counts = plan.run(planes)
            if "fitness" in what:
                device.population_fitness(counts, lut, out=out)
        torch.cuda.synchronize()
        print("fitness checksum", int(out.sum().item()) if "fitness" in what else int(counts.sum().item()))
    if "sweep" in what:
        net_parents, net_luts = compile_network(nodes)
        for _ in range(max(1, args.reps // 3)):
            res = device.ko_ki_sweep(planes, net_parents, net_luts, steps=50, stats=True)
        torch.cuda.synchronize()
        print("sweep stats", device.LAST_SWEEP_STATS)
        print("sweep checksum", int(res[0].sum().item()), "missing normal", int(res[1][0].item()))


if __name__ == "__main__":
    main()
