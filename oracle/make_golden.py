"""Generate ``tests/golden/*.json`` by running the UNMODIFIED scBONITA2 reference.

Run in the build container only (``python -m oracle.make_golden``): it imports the reference
from ``/root/reference/code`` through ``oracle/ref_loader.py``.  The committed JSON files carry
inputs AND reference outputs, so the parity tests need neither this script nor the reference.

Cases
-----
templates.json   ``Node.possibilities`` for 0..4 parents               (node_class.py:45-107)
rules_*.json     chunked matrix, ordered candidate lists, (difference, count) per candidate,
                 ``infer_ruleset()`` result and node side effects   (rule_determination.py)
sweep_*.json     seeded ``CalculateImportanceScore`` runs: sampled columns, RAW integer totals
                 (captured by shadowing ``max`` inside the reference module — the reference
                 only returns the normalised values), normalised scores, attractor lengths per
                 condition and cell                                  (importance_scores.py)
traj_*.json      ``attractor_analysis.vectorized_run_simulation`` for selected cells
"""
from __future__ import annotations

import builtins
import json
import logging
import os
import pickle
import random
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_loader  # noqa: E402
from tools import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rows_to_text(mat) -> list:
    return ["".join("1" if v else "0" for v in row) for row in np.asarray(mat)]


def spec_to_json(s) -> dict:
    return {"name": s.name, "index": s.index,
            "predecessors": [[int(k), v] for k, v in s.predecessors.items()],
            "inversions": [[int(k), bool(v)] for k, v in s.inversions.items()],
            "logic": s.logic}


def make_ref_nodes(ns, specs):
    return [ns.node_class.Node(s.name, s.index, dict(s.predecessors), dict(s.inversions))
            for s in specs]


def dump(name, obj):
    path = os.path.join(OUT, name)
    with open(path, "w") as fh:
        json.dump(obj, fh, separators=(",", ":"))
    print(f"wrote {path} ({os.path.getsize(path)} bytes)")


# ------------------------------------------------------------------------------------------
def golden_templates(ns):
    out = {}
    for k in range(5):
        node = ns.node_class.Node("X", 9, {i: f"P{i}" for i in range(k)},
                                  {i: False for i in range(k)})
        out[str(k)] = [str(x) for x in node.possibilities]
    dump("templates.json", out)


def golden_rules(ns, tag, n_nodes, n_cells, seed, flip, **net_kw):
    specs, planes = synth.synthetic_dataset(n_nodes, n_cells, seed, flip=flip, **net_kw)
    dense = synth.planes_to_dense(planes, n_cells)
    csr = sp.csr_matrix(dense.astype(float))
    RD = ns.rule_determination.RuleDetermination

    # candidates + per-candidate errors on fresh nodes (calculate_refined_errors mutates them)
    probe_nodes = make_ref_nodes(ns, specs)
    rd = RD(None, f"net_{tag}", f"ds_{tag}", csr, probe_nodes, {s.name: s.index for s in specs})
    chunked = np.asarray(rd.chunked_data_numpy)
    cands, errors = [], []
    for node in probe_nodes:
        node.find_all_rule_predictions()
    for node in probe_nodes:
        # same list growth as calculate_refined_errors (:459-475), by calling its building blocks
        rules = node.node_rules
        combos = [RD.generate_not_combinations(r[2]) for r in rules]
        for combo in combos:
            for logic in combo:
                parents = [p for p in node.predecessors]
                if [node.name, parents, logic] not in rules:
                    rules.append([node.name, parents, logic])
        cands.append([r[2] for r in rules])
        errors.append([[int(d), int(c)] for d, c in
                       (RD.calculate_error(node, r[2], chunked) for r in rules)])

    # the end-to-end inference on another fresh set of nodes
    nodes = make_ref_nodes(ns, specs)
    rd2 = RD(None, f"net_{tag}", f"ds_{tag}", csr, nodes, {s.name: s.index for s in specs})
    ruleset = rd2.infer_ruleset()
    out_rules = [[r[0][0], [int(i) for i in r[0][1]], r[0][2], int(r[1]), float(r[2])]
                 for r in ruleset]
    after = [{"predecessors": [[int(k), v] for k, v in n.predecessors.items()],
              "inversions": [[int(k), bool(v)] for k, v in n.inversions.items()],
              "n_node_rules": len(n.node_rules),
              "flat_node_rules": not isinstance(n.node_rules[0], list),
              "best_rule_index": int(n.best_rule_index)} for n in nodes]
    rules_txt = open(os.path.join(
        ns.file_paths.file_paths["rules_output"], f"ds_{tag}_rules",
        f"net_{tag}_ds_{tag}_rules.txt")).read()
    dump(f"rules_{tag}.json", {
        "n_cells": n_cells, "specs": [spec_to_json(s) for s in specs],
        "matrix": rows_to_text(dense), "num_chunks": int(rd.num_chunks),
        "chunked": rows_to_text(chunked), "candidates": cands, "errors": errors,
        "ruleset": out_rules, "nodes_after": after, "rules_txt": rules_txt})


def golden_chunk_widths(ns):
    """Chunk widths whose ties round differently: drive chunk_data directly."""
    RD = ns.rule_determination.RuleDetermination
    out = []
    rng = np.random.default_rng(5)
    for n_cells, num_chunks in [(240, 20), (280, 20), (300, 15), (247, 19), (100, 10),
                                (90, 10), (64, 64), (50, 80), (2095, 210)]:
        dense = (rng.random((5, n_cells)) < 0.5).astype(np.uint8)
        rd = RD.__new__(RD)
        rd.binarized_matrix = sp.csr_matrix(dense.astype(float))
        chunked = np.asarray(rd.chunk_data(num_chunks=num_chunks))
        out.append({"n_cells": n_cells, "num_chunks": num_chunks, "matrix": rows_to_text(dense),
                    "chunked": rows_to_text(chunked)})
    dump("chunk_widths.json", out)


class _CaptureMax:
    def __init__(self):
        self.values = None

    def __call__(self, *args, **kw):
        if len(args) == 1 and not kw:
            seq = list(args[0])
            self.values = seq
            return builtins.max(seq)
        return builtins.max(*args, **kw)


def golden_sweep(ns, tag, specs, dense, seed, logics=None):
    """``specs[i].logic`` (or ``logics[i]``) becomes ``best_rule[2]`` of node i."""
    nodes = make_ref_nodes(ns, specs)
    for i, node in enumerate(nodes):
        logic = logics[i] if logics else specs[i].logic
        node.best_rule = [node.name, [p for p in node.predecessors], logic]
    csr = sp.csr_matrix(dense.astype(float))
    random.seed(seed)
    mod = ns.importance_scores
    calc = mod.CalculateImportanceScore(nodes, csr.astype(bool), f"net_{tag}", f"ds_{tag}")
    cap = _CaptureMax()
    mod.max = cap
    result = {"n_cells": int(dense.shape[1]), "seed": seed,
              "specs": [spec_to_json(s) for s in specs],
              "logics": [n.best_rule[2] for n in nodes], "matrix": rows_to_text(dense),
              "sample": [int(i) for i in calc.cell_sample_indices]}
    try:
        scores = calc.calculate_importance_scores()
        result["raw"] = [int(v) for v in cap.values]
        result["scores"] = [float(scores[n.name]) for n in nodes]
        result["error"] = None
    except (KeyError, ZeroDivisionError) as exc:
        result["raw"] = [int(v) for v in cap.values] if cap.values is not None else None
        result["scores"] = None
        result["error"] = type(exc).__name__
    finally:
        del mod.max
    # attractor lengths (lambda+1; 0 = no repeat) per condition, from the reference's pickles
    n_cells = len(calc.cell_sample_indices)

    def lengths(path):
        found = pickle.load(open(path, "rb"))
        return [int(np.asarray(found[c]).shape[1]) if c in found else 0 for c in range(n_cells)]

    d = calc.intermediate_dir
    result["len_normal"] = lengths(f"{d}/normal_signaling.pkl")
    result["len_ko"] = [lengths(f"{d}/knockout_results_{n.name}.pkl") for n in nodes]
    result["len_ki"] = [lengths(f"{d}/knockin_results_{n.name}.pkl") for n in nodes]
    dump(f"sweep_{tag}.json", result)


def golden_traj(ns, tag, specs, dense, cells):
    nodes = make_ref_nodes(ns, specs)
    for node, s in zip(nodes, specs):
        node.calculation_function = s.logic
    out = {"specs": [spec_to_json(s) for s in specs], "matrix": rows_to_text(dense),
           "cells": cells,
           "trajectories": [ns.attractor_analysis.vectorized_run_simulation(
               nodes, [int(v) for v in dense[:, c]]) for c in cells]}
    dump(f"traj_{tag}.json", out)


def golden_rule_inference(ns, tag, n_genes, n_cells, seed, extra_rows=3):
    """The whole ``RuleInference`` constructor: CSV -> sampled cells -> binarised matrix ->
    Spearman top-3 parents -> nodes -> ruleset (rule_inference.py:21-101)."""
    import networkx as nx
    import tempfile
    rng = np.random.default_rng(seed)
    specs, planes = synth.synthetic_dataset(n_genes, n_cells, seed, flip=0.08, relax_steps=3)
    dense = synth.planes_to_dense(planes, n_cells)
    # expression values: 0/1 pattern times a positive magnitude, plus sub-threshold noise
    values = dense * rng.uniform(0.5, 3.0, size=dense.shape) + rng.uniform(0, 0.0005, size=dense.shape)
    names = [f"GENE{i}" for i in range(n_genes)]
    # a graph with up to 5 incoming edges per node so the top-3 pruning matters; node order == CSV order
    graph = nx.DiGraph()
    graph.add_nodes_from(names)
    edges = []
    for s in specs:
        parents = set(s.predecessors)
        while len(parents) < min(5, n_genes - 1) and rng.random() < 0.7:
            cand = int(rng.integers(0, n_genes))
            if cand != s.index:
                parents.add(cand)
        for p in sorted(parents):
            kind = "i" if (p in s.inversions and s.inversions[p]) or rng.random() < 0.1 else "a"
            key = "interaction" if rng.random() < 0.5 else "signal"
            graph.add_edge(names[p], names[s.index], **{key: kind})
            edges.append([names[p], names[s.index], key, kind])
    lines = ["gene," + ",".join(f"cell{c}" for c in range(n_cells))]
    for i in range(n_genes):
        lines.append(names[i] + "," + ",".join(f"{v:.6f}" for v in values[i]))
    for j in range(extra_rows):   # rows that are not part of the network
        lines.append(f"OTHER{j}," + ",".join(f"{v:.6f}" for v in rng.uniform(0, 2, n_cells)))
    csv_text = "\n".join(lines) + "\n"
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as fh:
        fh.write(csv_text)
        path = fh.name
    np.random.seed(seed)
    ri = ns.rule_inference.RuleInference(path, graph, f"ds_{tag}", f"net_{tag}", ",", list(range(n_genes)),
                                         binarize_threshold=0.001, sample_cells=True)
    os.unlink(path)
    dump(f"ruleinf_{tag}.json", {
        "seed": seed, "csv": csv_text, "node_names": names, "edges": edges,
        "node_indices": list(range(n_genes)),
        "gene_names": list(ri.gene_names), "cv_genes": list(ri.cv_genes),
        "binarized": rows_to_text(np.asarray(ri.binarized_matrix.todense())),
        "predecessors": [[int(x) for x in p] for p in ri.predecessors],
        "nodes": [{"predecessors": [[int(k), v] for k, v in n.predecessors.items()],
                   "inversions": [[int(k), bool(v)] for k, v in n.inversions.items()],
                   "best_rule": [n.best_rule[0], [int(i) for i in n.best_rule[1]], n.best_rule[2]],
                   "best_rule_index": int(n.best_rule_index)} for n in ri.nodes],
        "ruleset": [[r[0][0], [int(i) for i in r[0][1]], r[0][2], int(r[1]), float(r[2])] for r in ri.ruleset]})


def golden_rule_check(tag, n_genes, n_cells, seed):
    """``network_simulation/check_rules_output.py`` (imported unmodified): get_rules + check_rules on a
    planted network's rules (with every operator shape the planted rules use, upper- and lower-case
    spellings, parentheses) and a noisy data set, for the full set of cells and two subsets."""
    import importlib.util
    import tempfile
    spec = importlib.util.spec_from_file_location(
        "ref_check_rules_output", "/root/reference/network_simulation/check_rules_output.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(seed)
    specs, planes = synth.synthetic_dataset(n_genes, n_cells, seed, flip=0.1, relax_steps=3, p_inhibit=0.3)
    dense = synth.planes_to_dense(planes, n_cells)
    lines = []
    for s in specs:
        names = [f"Gene{p}" for p in s.predecessors] or [f"Gene{s.index}"]
        text = s.logic
        for slot, name in zip("ABC", names):
            text = text.replace(slot, "{" + slot + "}")
        text = text.format(**{slot: name for slot, name in zip("ABC", names)})
        if rng.random() < 0.5:
            text = text.replace(" and ", " AND ").replace(" or ", " OR ").replace("not ", "NOT ")
        lines.append(f"Gene{s.index} = {text}")
    rules_text = "\n".join(lines) + "\n"
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as fh:
        fh.write(rules_text)
        path = fh.name
    ruleset, raw = mod.get_rules(path)
    os.unlink(path)
    subsets = [list(range(n_cells)), sorted(int(c) for c in rng.choice(n_cells, int(0.8 * n_cells), replace=False)),
               [int(c) for c in rng.choice(n_cells, 7, replace=False)]]
    results = []
    for cells in subsets:
        m, t, e, k = mod.check_rules(dense, ruleset, raw, cells)
        results.append({"cells": cells, "mismatches": int(m), "total": int(t), "error_pct": float(e), "num_rules": int(k)})
    singles = [[int(mod.check_logic(rule, dense[:, c])) for rule in ruleset] for c in range(min(n_cells, 16))]
    dump(f"rulecheck_{tag}.json", {"n_genes": n_genes, "n_cells": n_cells, "rules_text": rules_text,
                                   "tokens": ruleset, "matrix": rows_to_text(dense), "results": results,
                                   "check_logic_first_cells": singles})


def seeded_graph(names, rng, max_in=5, p_more=0.7, specs=None):
    """A digraph over ``names`` (node order == name order) with up to ``max_in`` incoming edges per node,
    ~10-20 % inhibitory, edge attribute spelled ``interaction`` or ``signal`` at random."""
    import networkx as nx
    graph = nx.DiGraph()
    graph.add_nodes_from(names)
    edges = []
    n = len(names)
    for i in range(n):
        parents = set(specs[i].predecessors) if specs else set()
        if not specs:
            for _ in range(int(rng.integers(1, 4))):
                cand = int(rng.integers(0, n))
                if cand != i:
                    parents.add(cand)
        while len(parents) < min(max_in, n - 1) and rng.random() < p_more:
            cand = int(rng.integers(0, n))
            if cand != i:
                parents.add(cand)
        for p in sorted(parents):
            planted_inhibit = bool(specs and p in specs[i].inversions and specs[i].inversions[p])
            kind = "i" if planted_inhibit or rng.random() < 0.15 else "a"
            key = "interaction" if rng.random() < 0.5 else "signal"
            graph.add_edge(names[p], names[i], **{key: kind})
            edges.append([names[p], names[i], key, kind])
    return graph, edges


def hex_rows(mat) -> list:
    """0/1 rows as hex strings of the packed bits (bit b of byte k = column 8k + b)."""
    return [np.packbits(np.asarray(row, dtype=np.uint8), bitorder="little").tobytes().hex() for row in np.asarray(mat)]


def golden_tutorial(ns):
    """BASELINE.json configs[0]: the reference's tutorial data set (``input/tutorial_data/tutorial_04370_data.csv``,
    59 genes x 3621 cells, binarised at 0.01 as ``bash_scripts/local_tutorial.sh`` does) through the
    UNMODIFIED ``RuleInference`` (sampling, binarisation, Spearman top-3, rule determination with chunking:
    3621 > 2000 cells) and ``CalculateImportanceScore``.  The KEGG graph of pathway 04370 cannot be fetched
    offline, so the graph is a seeded digraph over the CSV's own 59 gene names (SURVEY 8-c fixture (i)).
    The fixture carries the binarised matrix in the reference's sampled cell order (not the CSV)."""
    csv_path = "/root/reference/input/tutorial_data/tutorial_04370_data.csv"
    with open(csv_path) as fh:
        next(fh)
        names = [line.split(",", 1)[0] for line in fh if line.strip()]
    rng = np.random.default_rng(4370)
    graph, edges = seeded_graph(names, rng)
    np.random.seed(4370)
    ri = ns.rule_inference.RuleInference(csv_path, graph, "ds_tutorial", "net_04370", ",", list(range(len(names))),
                                         binarize_threshold=0.01, sample_cells=True)
    binar = np.asarray(ri.binarized_matrix.todense()).astype(np.uint8)
    random.seed(4370)
    mod = ns.importance_scores
    calc = mod.CalculateImportanceScore(ri.nodes, ri.binarized_matrix.astype(bool), "net_04370", "ds_tutorial")
    cap = _CaptureMax()
    mod.max = cap
    try:
        scores = calc.calculate_importance_scores()
    finally:
        del mod.max
    dump("tutorial_04370.json", {
        "source": "input/tutorial_data/tutorial_04370_data.csv, threshold 0.01, np.random.seed(4370), random.seed(4370)",
        "gene_names": list(ri.gene_names), "n_cells": int(binar.shape[1]), "edges": edges,
        "binarized_hex": hex_rows(binar), "density": float(binar.mean()),
        "predecessors": [[int(x) for x in p] for p in ri.predecessors],
        "nodes": [{"predecessors": [[int(k), v] for k, v in n.predecessors.items()],
                   "inversions": [[int(k), bool(v)] for k, v in n.inversions.items()],
                   "best_rule": [n.best_rule[0], [int(i) for i in n.best_rule[1]], n.best_rule[2]],
                   "best_rule_index": int(n.best_rule_index)} for n in ri.nodes],
        "ruleset": [[r[0][0], [int(i) for i in r[0][1]], r[0][2], int(r[1]), float(r[2])] for r in ri.ruleset],
        "sample": [int(i) for i in calc.cell_sample_indices],
        "raw": [int(v) for v in cap.values], "scores": [float(scores[n.name]) for n in ri.nodes]})


def golden_pickles(ns):
    """Files WRITTEN BY THE REFERENCE for the interchange tests: its own
    ``Pipeline.infer_rules`` (``pipeline_class.py:134-182``) writes ``*.ruleset.pickle`` and ``*.cells.pickle``;
    its own ``run_full_importance_score`` (``importance_scores.py:343-427``) reads the former and writes the
    importance text file and ``*.network.pickle``.  Copied byte for byte into tests/golden/pickles/."""
    import importlib
    import shutil
    import tempfile
    sys.modules.pop("pipeline_class", None)
    pipeline_class = importlib.import_module("pipeline_class")
    rng = np.random.default_rng(81)
    n_genes, n_cells, seed = 9, 60, 81
    specs, planes = synth.synthetic_dataset(n_genes, n_cells, seed, flip=0.08, relax_steps=3)
    dense = synth.planes_to_dense(planes, n_cells)
    values = dense * rng.uniform(0.5, 3.0, size=dense.shape)
    names = [f"GENE{i}" for i in range(n_genes)]
    graph, edges = seeded_graph(names, rng, max_in=4, p_more=0.5, specs=specs)
    lines = ["gene," + ",".join(f"cell{c}" for c in range(n_cells))]
    lines += [names[i] + "," + ",".join(f"{v:.6f}" for v in values[i]) for i in range(n_genes)]
    csv_text = "\n".join(lines) + "\n"
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as fh:
        fh.write(csv_text)
        csv_path = fh.name
    pipe = pipeline_class.Pipeline.__new__(pipeline_class.Pipeline)   # __init__ parses argv and fetches KEGG
    pipe.data_file, pipe.dataset_name, pipe.datafile_sep = csv_path, "ds_pickle", ","
    pipe.binarize_threshold, pipe.sample_cells, pipe.organism = 0.001, True, "hsa"
    for path in ns.file_paths.file_paths.values():
        os.makedirs(path, exist_ok=True)
    np.random.seed(seed)
    pipe.infer_rules("00081", graph, list(range(n_genes)))
    os.unlink(csv_path)
    random.seed(seed)
    # the importance figure needs a real matplotlib, which this container lacks: networkx's three drawing
    # calls (not reference code) become no-ops for the duration of the call
    import networkx as nx
    saved = {name: getattr(nx, name) for name in ("draw_networkx_nodes", "draw_networkx_edges", "draw_networkx_labels")}
    for name in saved:
        setattr(nx, name, lambda *a, **k: None)
    try:
        ns.importance_scores.run_full_importance_score("ds_pickle", ["hsa00081"])
    finally:
        for name, fn in saved.items():
            setattr(nx, name, fn)
    fp = ns.file_paths.file_paths
    out = os.path.join(OUT, "pickles")
    os.makedirs(out, exist_ok=True)
    base = f'{fp["pickle_files"]}/ds_pickle_pickle_files'
    copies = {f"{base}/ruleset_pickle_files/ds_pickle_hsa00081.ruleset.pickle": "ds_pickle_hsa00081.ruleset.pickle",
              f"{base}/cells_pickle_file/ds_pickle.cells.pickle": "ds_pickle.cells.pickle",
              f"{base}/network_pickle_files/ds_pickle_hsa00081.network.pickle": "ds_pickle_hsa00081.network.pickle",
              f'{fp["importance_score_output"]}/ds_pickle/text_files/hsa00081_importance_score.txt': "hsa00081_importance_score.txt",
              f'{fp["rules_output"]}/ds_pickle_rules/00081_ds_pickle_rules.txt': "00081_ds_pickle_rules.txt"}
    for src, dst in copies.items():
        shutil.copyfile(src, os.path.join(out, dst))
        print(f"wrote {os.path.join(out, dst)} ({os.path.getsize(src)} bytes)")
    dump("pickles/inputs.json", {"seed": seed, "csv": csv_text, "node_names": names, "edges": edges,
                                 "dataset_name": "ds_pickle", "pathway": "00081", "organism": "hsa"})


def ring_network(n, twist=True):
    """Shift ring G_i <- G_{i-1}; with ``twist`` G_0 <- not G_{n-1} (period 2n Johnson ring)."""
    specs = []
    for i in range(n):
        p = (i - 1) % n
        inv = bool(twist and i == 0)
        specs.append(synth.NodeSpec(f"G{i}", i, {p: f"G{p}"}, {p: inv},
                                    "not A" if inv else "A"))
    return specs


def main():
    logging.disable(logging.CRITICAL)
    os.makedirs(OUT, exist_ok=True)
    ns = ref_loader.load()
    golden_templates(ns)
    golden_chunk_widths(ns)

    # rule inference: small unchunked, noisy, and chunked (>2000 cells; widths 10 and 9)
    golden_rules(ns, "small", 12, 150, seed=1, flip=0.05)
    golden_rules(ns, "noisy", 16, 97, seed=2, flip=0.25)
    golden_rules(ns, "selfloops", 10, 64, seed=3, flip=0.10, p_source=0.4, allow_self=True)
    golden_rules(ns, "chunk10", 8, 2250, seed=4, flip=0.10)
    golden_rules(ns, "chunk9", 6, 2095, seed=5, flip=0.15)

    # KO/KI sweeps: planted random networks (mostly fixed points / short cycles) ...
    for tag, n_nodes, n_cells, seed in [("rand_a", 10, 40, 11), ("rand_b", 14, 70, 12),
                                        ("rand_c", 20, 33, 13)]:
        specs, planes = synth.synthetic_dataset(n_nodes, n_cells, seed, flip=0.2, relax_steps=2)
        golden_sweep(ns, tag, specs, synth.planes_to_dense(planes, n_cells), seed)
    # ... more cells than the 500-cell sample
    specs, planes = synth.synthetic_dataset(6, 620, 21, flip=0.2, relax_steps=1)
    golden_sweep(ns, "sampled", specs, synth.planes_to_dense(planes, 620), 21)
    # ... long cycles: plain 5-ring (period | 5), Johnson rings with period 12 and 2*30 > 50
    rng = np.random.default_rng(31)
    for tag, n, twist in [("ring5", 5, False), ("twist6", 6, True), ("twist30", 30, True)]:
        specs = ring_network(n, twist)
        dense = (rng.random((n, 24)) < 0.5).astype(np.uint8)
        golden_sweep(ns, tag, specs, dense, 31)
    # ... mixed: a random network feeding from a 2-cycle and a 3-ring
    specs, planes = synth.synthetic_dataset(9, 48, 41, flip=0.3, relax_steps=0)
    logics = [s.logic for s in specs]
    specs[0].predecessors, specs[0].inversions, logics[0] = {1: "G1"}, {1: True}, "not A"
    specs[1].predecessors, specs[1].inversions, logics[1] = {0: "G0"}, {0: False}, "A"
    specs[2].predecessors, specs[2].inversions, logics[2] = {4: "G4"}, {4: False}, "A"
    specs[3].predecessors, specs[3].inversions, logics[3] = {2: "G2"}, {2: False}, "A"
    specs[4].predecessors, specs[4].inversions, logics[4] = {3: "G3"}, {3: True}, "not A"
    for s, lg in zip(specs, logics):
        s.logic = lg
    golden_sweep(ns, "mixed", specs, synth.planes_to_dense(planes, 48), 41)

    golden_tutorial(ns)
    golden_pickles(ns)
    golden_rule_check("a", 12, 90, seed=71)
    golden_rule_check("b", 30, 257, seed=72)

    golden_rule_inference(ns, "a", 10, 80, seed=61)
    golden_rule_inference(ns, "b", 14, 2100, seed=62)

    # trajectories
    specs, planes = synth.synthetic_dataset(12, 30, 51, flip=0.3, relax_steps=0)
    golden_traj(ns, "rand", specs, synth.planes_to_dense(planes, 30), [0, 3, 7, 29])
    specs = ring_network(7, True)
    golden_traj(ns, "twist7", specs, (rng.random((7, 6)) < 0.5).astype(np.uint8), [0, 1, 5])


if __name__ == "__main__":
    main()
