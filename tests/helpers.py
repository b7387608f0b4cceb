"""Shared helpers for the parity tests (golden loading, node construction)."""
import json
import os

import numpy as np

from scbonita2_b200.node_class import Node

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def golden_names(prefix):
    return sorted(f for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".json"))


def matrix_of(rows_text) -> np.ndarray:
    return np.array([[1 if ch == "1" else 0 for ch in row] for row in rows_text], dtype=np.uint8)


def nodes_of(specs, with_logic=False):
    """Product Node objects from golden / synth specs (dicts or NodeSpec)."""
    nodes = []
    for s in specs:
        if isinstance(s, dict):
            preds = {int(k): v for k, v in s["predecessors"]}
            invs = {int(k): bool(v) for k, v in s["inversions"]}
            node = Node(s["name"], s["index"], preds, invs)
            logic = s["logic"]
        else:
            node = Node(s.name, s.index, dict(s.predecessors), dict(s.inversions))
            logic = s.logic
        if with_logic:
            node.best_rule = [node.name, [p for p in node.predecessors], logic]
            node.calculation_function = logic
        nodes.append(node)
    return nodes


def np_counts(dense, target, parents):
    """Reference joint counts [16] for one group by plain NumPy."""
    t = dense[target].astype(np.int64)
    cols = [dense[p].astype(np.int64) if p >= 0 else np.zeros_like(t) for p in parents]
    k = cols[0] * 4 + cols[1] * 2 + cols[2]
    out = np.zeros(16, dtype=np.int64)
    np.add.at(out, 2 * k + t, 1)
    return out


def np_diff(counts16, lut):
    return int(sum(counts16[2 * k + (1 - ((lut >> k) & 1))] for k in range(8)))
