"""BASELINE.json configs[4]: a batch of KEGG-sized pathways x 500k cells, sharded by pathway across
the ranks of one box (pathways are independent: no collective at all), cells kept whole per pathway.

    python tools/run_k50.py [--pathways 50] [--cells 500000]            # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/run_k50.py --pathways 50

Per pathway: pack -> joint counts -> error table of every candidate rule (rule inference's GPU
part) -> KO/KI importance sweep.  Prints one JSON line with per-stage and total times.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tools import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pathways", type=int, default=50)
    ap.add_argument("--cells", type=int, default=500_000)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench
    from scbonita2_b200 import device, rule_compiler
    from scbonita2_b200.importance_scores import compile_network

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    rng = np.random.default_rng(args.seed)
    sizes = rng.integers(30, 201, args.pathways)          # N ~ U{30..200}
    mine = [p for p in range(args.pathways) if p % world == rank]
    # inputs for my pathways (host side, untimed)
    inputs = []
    for p in mine:
        specs = synth.random_network(int(sizes[p]), args.seed + p)
        start = synth.random_planes(int(sizes[p]), args.cells, args.seed + 1000 + p)
        planes_np = synth.relax_and_flip(specs, start, args.cells, args.seed + p, relax_steps=10, flip=0.05)
        inputs.append((p, specs, planes_np))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_all = time.perf_counter()
    t_rules = t_sweep = 0.0
    checks = {}
    for p, specs, planes_np in inputs:
        nodes = bench.make_nodes(specs)
        t0 = time.perf_counter()
        planes = device.BitPlanes.from_packed(planes_np, args.cells)
        targets, parents, group, lut = [], [], [], []
        for g, node in enumerate(nodes):
            node.find_all_rule_predictions()
            rules = rule_compiler.extend_with_not_variants(node)
            targets.append(node.index)
            parents.append(rule_compiler.parent_rows(node))
            slots = min(3, len(node.predecessors))
            for r in rules:
                group.append(g)
                lut.append(rule_compiler.compile_for_slots(r[2], slots))
        counts = device.rule_counts(planes, targets, parents)
        diffs = device.rule_error_table(counts, group, lut)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        net_parents, net_luts = compile_network(nodes)
        raw, missing, keyerr = device.ko_ki_sweep(planes, net_parents, net_luts)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t_rules += t1 - t0
        t_sweep += t2 - t1
        checks[p] = [int(diffs.sum().item()), int(raw.sum().item())]
    total = time.perf_counter() - t_all
    if world > 1:
        t = torch.tensor([total, t_rules, t_sweep], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total, t_rules, t_sweep = t.tolist()
        gathered = [None] * world
        dist.all_gather_object(gathered, checks)
        checks = {k: v for d in gathered for k, v in d.items()}
    if rank == 0:
        print(json.dumps({"config": "K50: batch of pathways x cells, sharded by pathway", "pathways": args.pathways,
                          "cells": args.cells, "n_gpus": world, "nodes_total": int(sizes.sum()),
                          "seconds_total": total, "seconds_rule_scoring": t_rules, "seconds_sweep": t_sweep,
                          "checksum": int(sum(a + b for a, b in checks.values()))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
