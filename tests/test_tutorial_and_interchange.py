"""cfg1 (the reference's tutorial data set) and file interchange with the reference.

``tests/golden/tutorial_04370.json`` and ``tests/golden/pickles/*`` were produced by the UNMODIFIED
reference (``oracle/make_golden.py``: ``golden_tutorial``, ``golden_pickles``).  CPU tests pin the oracle
ports and the alias loading; GPU tests replay the product's entry points on the same inputs."""
import csv
import json
import os
import pickle
import random
import shutil
import subprocess
import sys

import networkx as nx
import numpy as np
import pytest

from tests.helpers import GOLDEN, golden

PICKLES = os.path.join(GOLDEN, "pickles")


def _tutorial():
    g = golden("tutorial_04370.json")
    n_cells = g["n_cells"]
    rows = [np.unpackbits(np.frombuffer(bytes.fromhex(h), dtype=np.uint8), bitorder="little")[:n_cells]
            for h in g["binarized_hex"]]
    return g, np.array(rows, dtype=np.uint8)


def _graph(names, edges):
    graph = nx.DiGraph()
    graph.add_nodes_from(names)
    for src, dst, key, kind in edges:
        graph.add_edge(src, dst, **{key: kind})
    return graph


def _write_csv(path, names, matrix):
    with open(path, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["gene"] + [f"cell{c}" for c in range(matrix.shape[1])])
        for name, row in zip(names, matrix):
            w.writerow([name] + [f"{float(v):.1f}" for v in row])


# ---------------------------------------------------------------- CPU: oracle ports on the real data
def test_tutorial_fixture_shape_and_density():
    g, dense = _tutorial()
    assert dense.shape == (59, 3621) and abs(dense.mean() - g["density"]) < 1e-12
    assert 0.2 < g["density"] < 0.26          # 0.23 at threshold 0.01


def test_oracle_ports_reproduce_the_tutorial_run():
    """Eval-per-cell port on the chunked tutorial matrix for the selected rules, and the C sweep port on
    the 500 sampled cells, against what the reference produced."""
    from oracle import cport, ref_port
    from scbonita2_b200.importance_scores import compile_network
    from scbonita2_b200.node_class import Node
    from tools import synth
    g, dense = _tutorial()
    nodes = []
    for i, rec in enumerate(g["nodes"]):
        node = Node(g["gene_names"][i], i, {int(k): v for k, v in rec["predecessors"]},
                    {int(k): bool(v) for k, v in rec["inversions"]})
        node.best_rule = list(rec["best_rule"])
        node.calculation_function = rec["best_rule"][2]
        nodes.append(node)
    chunked = ref_port.chunk_data(dense, round(dense.shape[1] / 10))
    for i in (0, 7, 23, 58):
        name, parents, logic = g["ruleset"][i][:3]
        if parents and parents[0] != i:
            diff, count = ref_port.calculate_error(nodes[i], logic, chunked.astype(float))
            assert diff / count == pytest.approx(g["ruleset"][i][4], abs=1e-12)
    parents, luts = compile_network(nodes)
    sample = dense[:, g["sample"]]
    raw, missing, keyerr = cport.importance_raw_bitsliced(synth.dense_to_planes(sample), sample.shape[1], parents, luts, 50)
    assert keyerr == 0 and raw.tolist() == g["raw"]
    mx = max(g["raw"])
    assert [r / mx for r in g["raw"]] == g["scores"]


def test_reference_written_pickles_load_under_the_aliases():
    """A6: ``*.ruleset.pickle`` / ``*.cells.pickle`` / ``*.network.pickle`` written by the reference's own
    Pipeline.infer_rules and run_full_importance_score unpickle into the product's classes."""
    code = r"""
import json, pickle, sys
sys.path.insert(0, %r)
import scbonita2_b200
scbonita2_b200.install_aliases()
from scbonita2_b200 import rule_inference, network_class, cell_class, node_class
d = %r
rs = pickle.load(open(d + '/ds_pickle_hsa00081.ruleset.pickle', 'rb'))
cp = pickle.load(open(d + '/ds_pickle.cells.pickle', 'rb'))
nw = pickle.load(open(d + '/ds_pickle_hsa00081.network.pickle', 'rb'))
assert type(rs) is rule_inference.RuleInference and type(nw) is network_class.Network
assert type(cp) is cell_class.CellPopulation and cp.cells == []          # SURVEY Q13
assert all(type(n) is node_class.Node for n in rs.nodes) and len(nw.nodes) == len(rs.nodes) == 9
print(json.dumps({'rules': [[r[0][0], list(map(int, r[0][1])), r[0][2]] for r in rs.ruleset],
                  'scores': [[n.name, round(float(n.importance_score), 3)] for n in nw.nodes],
                  'shape': list(rs.binarized_matrix.shape)}))
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), PICKLES)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    got = json.loads(out.strip().splitlines()[-1])
    assert got["shape"] == [9, 60]
    want = [line.split("\t") for line in open(os.path.join(PICKLES, "hsa00081_importance_score.txt")).read().splitlines()[1:]]
    assert [[n, f"{s}"] for n, s in got["scores"]] == [[n, str(float(s))] for n, s in want]
    rules_txt = open(os.path.join(PICKLES, "00081_ds_pickle_rules.txt")).read().splitlines()
    assert len(rules_txt) == len(got["rules"]) == 9


# ---------------------------------------------------------------- GPU: the product on the same inputs
@pytest.mark.gpu
@pytest.mark.parametrize("gpu_pruning", [False, True])
def test_tutorial_run_matches_the_reference(cuda, tmp_path, gpu_pruning):
    """RuleInference (ingest -> binarise -> Spearman top-3 -> chunked rule determination on the GPU) and
    CalculateImportanceScore on the tutorial matrix: parents, inversions, selected rules, errors, sampled
    cells, raw totals and normalised scores equal the reference's."""
    import logging
    from scbonita2_b200 import file_paths
    from scbonita2_b200.importance_scores import CalculateImportanceScore
    from scbonita2_b200.rule_inference import RuleInference
    logging.disable(logging.CRITICAL)
    file_paths.set_output_root(str(tmp_path / "out"))
    g, dense = _tutorial()
    path = str(tmp_path / "tutorial01.csv")
    _write_csv(path, g["gene_names"], dense)       # the reference's sampled order, already binarised
    ri = RuleInference(path, _graph(g["gene_names"], g["edges"]), "ds_tutorial", "net_04370", ",",
                       list(range(59)), binarize_threshold=0.01, sample_cells=False, gpu_parent_pruning=gpu_pruning)
    assert (np.asarray(ri.binarized_matrix.todense()).astype(np.uint8) == dense).all()
    assert [[int(x) for x in p] for p in ri.predecessors] == g["predecessors"]
    for node, want in zip(ri.nodes, g["nodes"]):
        assert [[int(k), v] for k, v in node.predecessors.items()] == want["predecessors"]
        assert [[int(k), bool(v)] for k, v in node.inversions.items()] == want["inversions"]
        assert [node.best_rule[0], [int(i) for i in node.best_rule[1]], node.best_rule[2]] == want["best_rule"]
        assert int(node.best_rule_index) == want["best_rule_index"]
    got = [[r[0][0], [int(i) for i in r[0][1]], r[0][2], int(r[1]), float(r[2])] for r in ri.ruleset]
    assert got == g["ruleset"]
    random.seed(4370)
    calc = CalculateImportanceScore(ri.nodes, ri.binarized_matrix.astype(bool), "net_04370", "ds_tutorial")
    assert calc.cell_sample_indices == g["sample"]
    scores = calc.calculate_importance_scores()
    assert [calc.raw_importance_scores[n.name] for n in ri.nodes] == g["raw"]
    assert [float(scores[n.name]) for n in ri.nodes] == g["scores"]


def _stage_reference_ruleset(root):
    from scbonita2_b200 import file_paths
    fp = file_paths.set_output_root(root)
    d = f'{fp["pickle_files"]}/ds_pickle_pickle_files/ruleset_pickle_files'
    os.makedirs(d, exist_ok=True)
    shutil.copyfile(os.path.join(PICKLES, "ds_pickle_hsa00081.ruleset.pickle"), f"{d}/ds_pickle_hsa00081.ruleset.pickle")
    return fp


@pytest.mark.gpu
def test_run_full_importance_score_consumes_a_reference_ruleset_pickle(cuda, tmp_path):
    """The reference-written ruleset pickle through the product's run_full_importance_score: the score
    file is byte-equal to the one the reference wrote, the network pickle carries the same scores."""
    import logging
    import scbonita2_b200
    from scbonita2_b200.importance_scores import run_full_importance_score
    logging.disable(logging.CRITICAL)
    scbonita2_b200.install_aliases()
    fp = _stage_reference_ruleset(str(tmp_path / "out"))
    inputs = golden("pickles/inputs.json")
    random.seed(inputs["seed"])
    run_full_importance_score("ds_pickle", ["hsa00081"])
    ours = open(f'{fp["importance_score_output"]}/ds_pickle/text_files/hsa00081_importance_score.txt').read()
    assert ours == open(os.path.join(PICKLES, "hsa00081_importance_score.txt")).read()
    with open(f'{fp["pickle_files"]}/ds_pickle_pickle_files/network_pickle_files/ds_pickle_hsa00081.network.pickle', "rb") as fh:
        mine = pickle.load(fh)
    with open(os.path.join(PICKLES, "ds_pickle_hsa00081.network.pickle"), "rb") as fh:
        theirs = pickle.load(fh)
    assert [float(n.importance_score) for n in mine.nodes] == [float(n.importance_score) for n in theirs.nodes]
    assert [n.name for n in mine.nodes] == [n.name for n in theirs.nodes]
    with pytest.raises(Exception):
        run_full_importance_score("ds_pickle", ["hsa99999"])      # as the reference: unknown pathway


@pytest.mark.gpu
def test_pipeline_infer_rules_writes_what_the_reference_writes(cuda, tmp_path):
    """pipeline_class.Pipeline.infer_rules on the inputs the reference's own Pipeline.infer_rules was run on
    (same NumPy seed): same rules file, same ruleset and parents as in the reference-written pickle, and
    pickles that name the reference's class paths."""
    import logging
    import pickletools
    import scbonita2_b200
    from scbonita2_b200 import file_paths
    from scbonita2_b200.pipeline_class import Pipeline
    logging.disable(logging.CRITICAL)
    scbonita2_b200.install_aliases()
    fp = file_paths.set_output_root(str(tmp_path / "out"))
    inputs = golden("pickles/inputs.json")
    path = str(tmp_path / "data.csv")
    open(path, "w").write(inputs["csv"])
    pipe = Pipeline(path, "ds_pickle", ",", 0.001, "hsa", True, minimum_overlap=1)
    np.random.seed(inputs["seed"])
    ours = pipe.infer_rules(inputs["pathway"], _graph(inputs["node_names"], inputs["edges"]), list(range(len(inputs["node_names"]))))
    with open(os.path.join(PICKLES, "ds_pickle_hsa00081.ruleset.pickle"), "rb") as fh:
        theirs = pickle.load(fh)
    assert ours.ruleset == theirs.ruleset and ours.predecessors == theirs.predecessors
    assert (ours.binarized_matrix != theirs.binarized_matrix).nnz == 0
    assert open(f'{fp["rules_output"]}/ds_pickle_rules/00081_ds_pickle_rules.txt').read() == \
        open(os.path.join(PICKLES, "00081_ds_pickle_rules.txt")).read()
    written = f'{fp["pickle_files"]}/ds_pickle_pickle_files/ruleset_pickle_files/ds_pickle_hsa00081.ruleset.pickle'
    names = {arg for op, arg, _ in pickletools.genops(open(written, "rb").read()) if op.name in ("GLOBAL", "STACK_GLOBAL", "SHORT_BINUNICODE")
             and isinstance(arg, str)}
    assert "rule_inference" in names and "node_class" in names and not any(n.startswith("scbonita2_b200") for n in names)
    cells = pickle.load(open(f'{fp["pickle_files"]}/ds_pickle_pickle_files/cells_pickle_file/ds_pickle.cells.pickle', "rb"))
    assert cells.cells == []
    # rule_inference(graphs): the per-pathway loop of the reference (pipeline_class.py:108-131)
    pipe2 = Pipeline(path, "ds_loop", ",", 0.001, "hsa", False, minimum_overlap=1)
    pipe2.rule_inference({"00081": _graph(inputs["node_names"], inputs["edges"])})
    assert os.path.exists(f'{fp["pickle_files"]}/ds_loop_pickle_files/ruleset_pickle_files/ds_loop_hsa00081.ruleset.pickle')


@pytest.mark.gpu
def test_the_reference_reads_our_ruleset_pickle(cuda, tmp_path):
    """The other direction: the UNMODIFIED reference (baseline/_ref or /root/reference, in a child process)
    unpickles the ruleset pickle the product wrote and runs its own run_full_importance_score on it; its
    score file equals the product's."""
    import logging
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("the reference sources are not installed on this box")
    import scbonita2_b200
    from scbonita2_b200 import file_paths
    from scbonita2_b200.importance_scores import run_full_importance_score
    from scbonita2_b200.pipeline_class import Pipeline
    logging.disable(logging.CRITICAL)
    scbonita2_b200.install_aliases()
    fp = file_paths.set_output_root(str(tmp_path / "out"))
    inputs = golden("pickles/inputs.json")
    path = str(tmp_path / "data.csv")
    open(path, "w").write(inputs["csv"])
    np.random.seed(5)
    Pipeline(path, "ds_ours", ",", 0.001, "hsa", True, minimum_overlap=1).infer_rules(
        "00081", _graph(inputs["node_names"], inputs["edges"]), list(range(len(inputs["node_names"]))))
    random.seed(9)
    run_full_importance_score("ds_ours", ["hsa00081"])
    ours = open(f'{fp["importance_score_output"]}/ds_ours/text_files/hsa00081_importance_score.txt').read()
    written = f'{fp["pickle_files"]}/ds_ours_pickle_files/ruleset_pickle_files/ds_ours_hsa00081.ruleset.pickle'
    code = r"""
import logging, os, random, shutil, sys
sys.path.insert(0, %r)
logging.disable(logging.CRITICAL)
from oracle import ref_loader
ns = ref_loader.load()
import networkx as nx
for name in ("draw_networkx_nodes", "draw_networkx_edges", "draw_networkx_labels"):
    setattr(nx, name, lambda *a, **k: None)      # no matplotlib in this image; the figure is off the path
fp = ns.file_paths.file_paths
d = fp["pickle_files"] + "/ds_ours_pickle_files/ruleset_pickle_files"
os.makedirs(d, exist_ok=True)
shutil.copyfile(%r, d + "/ds_ours_hsa00081.ruleset.pickle")
random.seed(9)
ns.importance_scores.run_full_importance_score("ds_ours", ["hsa00081"])
print("SCORES" + open(fp["importance_score_output"] + "/ds_ours/text_files/hsa00081_importance_score.txt").read())
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), written)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert proc.returncode == 0, proc.stderr[-2000:]
    theirs = proc.stdout.split("SCORES", 1)[1].rstrip("\n")
    assert theirs == ours.rstrip("\n")
