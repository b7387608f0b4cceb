"""BASELINE.json configs[1] ("S100k"): rule-recovery accuracy on a synthetic Boolean network with known
rules, 100k simulated cells, one B200.

    python tools/run_s100k.py [--cells 100000] [--sizes 10,20,50,100] [--reps 5]

Per network size N: planted network (``tools/synth.py``, the shape of the reference's
``network_simulation/create_test_network.py``) -> cells relaxed under the planted rules + 10 % flips (the
reference's mutation rate, ``network_simulation.py:245-252``) -> ``RuleDetermination.infer_ruleset()`` on the
GPU -> the inferred and the planted rules are both scored against the data the reference's way
(``check_rules_output.py``: error_pct over 80 % cell subsamples; ``scbonita2_b200.rule_recovery`` computes
it on the GPU), next to the share of nodes whose inferred truth table equals the planted one and to
the GPU time of the inference.  Prints one JSON line per size and a summary line.
"""
import argparse
import json
import logging
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tools import synth  # noqa: E402


def rule_line(name, logic, parent_names):
    text = logic
    for slot in "ABC":
        text = text.replace(slot, "{" + slot + "}")
    return f"{name} = " + text.format(**{slot: p for slot, p in zip("ABC", parent_names)})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100_000)
    ap.add_argument("--sizes", default="10,20,50,100")
    ap.add_argument("--reps", type=int, default=5, help="80 %% subsamples per size (the reference uses 50)")
    ap.add_argument("--flip", type=float, default=0.10)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import torch

    from scbonita2_b200 import device, file_paths, rule_compiler, rule_recovery
    from scbonita2_b200.node_class import Node
    from scbonita2_b200.rule_determination import RuleDetermination
    logging.disable(logging.CRITICAL)
    file_paths.set_output_root(os.environ.get("SCBONITA_OUTPUT", "/tmp/scbonita_s100k"))
    summary = []
    for n in [int(x) for x in args.sizes.split(",")]:
        specs = synth.random_network(n, args.seed + n)
        for s in specs:
            s.name = f"Gene{s.index}"
            s.predecessors = {p: f"Gene{p}" for p in s.predecessors}
        start = synth.random_planes(n, args.cells, args.seed + 1000 + n)
        planes_np = synth.relax_and_flip(specs, start, args.cells, args.seed + n, relax_steps=30, flip=args.flip)
        nodes = [Node(s.name, s.index, dict(s.predecessors), dict(s.inversions)) for s in specs]
        planes = device.BitPlanes.from_packed(planes_np, args.cells)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rd = RuleDetermination(None, f"net{n}", f"s100k_{n}", planes, nodes, {s.name: s.index for s in specs})
        ruleset = rd.infer_ruleset()
        torch.cuda.synchronize()
        t_infer = time.perf_counter() - t0
        # rule texts the checker reads: "GeneX = GeneA and not GeneB"
        name_of = {s.index: s.name for s in specs}
        inferred = "\n".join(rule_line(r[0][0], r[0][2], [name_of[i] for i in r[0][1]]) for r in ruleset) + "\n"
        planted = "\n".join(rule_line(s.name, s.logic, list(s.predecessors.values()) or [s.name]) for s in specs) + "\n"
        inf_rules, inf_raw = rule_recovery.get_rules(inferred)
        pla_rules, pla_raw = rule_recovery.get_rules(planted)
        rng = np.random.default_rng(args.seed + 7 * n)
        err_inf, err_pla = [], []
        t0 = time.perf_counter()
        for _ in range(args.reps):
            cells = rng.choice(args.cells, size=max(1, int(args.cells * 0.8)), replace=False)
            err_inf.append(rule_recovery.check_rules(planes, inf_rules, inf_raw, cells)[2])
            err_pla.append(rule_recovery.check_rules(planes, pla_rules, pla_raw, cells)[2])
        t_check = time.perf_counter() - t0
        # exact recovery: same truth table over the same parent rows as the planted rule
        same = 0
        for s, r in zip(specs, ruleset):
            rows = list(r[0][1])[:3]
            lut = rule_compiler.compile_for_slots(r[0][2], min(3, len(rows)))
            want_rows = list(s.predecessors)[:3] or [s.index]
            same += int(rows == want_rows and lut == rule_compiler.compile_for_slots(s.logic, len(want_rows)))
        line = {"config": "S100k rule recovery", "nodes": n, "cells": args.cells, "flip": args.flip,
                "error_pct_inferred": float(np.mean(err_inf)), "error_pct_planted": float(np.mean(err_pla)),
                "rules_identical_to_planted": same, "seconds_infer_ruleset": t_infer,
                "seconds_checker_per_subsample": t_check / (2 * args.reps),
                "candidate_rule_evals_per_s": sum(len(nd.node_rules) for nd in nodes) * float(rd.chunked_planes.n_cells) / t_infer}
        print(json.dumps(line), flush=True)
        summary.append(line)
    print(json.dumps({"config": "S100k summary", "sizes": [s["nodes"] for s in summary],
                      "error_pct_inferred": [round(s["error_pct_inferred"], 3) for s in summary],
                      "error_pct_planted": [round(s["error_pct_planted"], 3) for s in summary],
                      "identical": [s["rules_identical_to_planted"] for s in summary]}))


if __name__ == "__main__":
    main()
