"""GPU parity: rule scoring (joint counts, error table, population fitness, inference)."""
import logging
import os

import numpy as np
import pytest

from oracle import ref_port
from scbonita2_b200 import rule_compiler
from tests.helpers import golden, golden_names, matrix_of, nodes_of, np_counts, np_diff
from tools import synth

pytestmark = pytest.mark.gpu


def _groups(specs):
    targets = [s.index for s in specs]
    parents = [synth.spec_parent_rows(s) for s in specs]
    return targets, parents


@pytest.mark.parametrize("n_nodes,n_cells,seed", [(1, 1, 0), (4, 31, 1), (5, 32, 2), (8, 33, 3),
                                                  (12, 1000, 4), (59, 3621, 5), (30, 20011, 6),
                                                  (200, 4099, 7)])
def test_counts_match_numpy(cuda, n_nodes, n_cells, seed):
    from scbonita2_b200 import device
    specs, planes_np = synth.synthetic_dataset(n_nodes, n_cells, seed, flip=0.1)
    dense = synth.planes_to_dense(planes_np, n_cells)
    planes = device.BitPlanes.from_packed(planes_np, n_cells)
    targets, parents = _groups(specs)
    counts = device.rule_counts(planes, targets, parents).cpu().numpy()
    for g in range(n_nodes):
        assert (counts[g] == np_counts(dense, targets[g], parents[g])).all(), g
        assert counts[g].sum() == n_cells


def test_counts_degenerate_groups(cuda):
    """Duplicate parents, parent == target, unused slots, more groups than rows."""
    from scbonita2_b200 import device
    rng = np.random.default_rng(3)
    dense = (rng.random((5, 777)) < 0.4).astype(np.uint8)
    planes = device.BitPlanes.from_dense(dense)
    targets = [0, 1, 2, 3, 4, 0, 0, 2]
    parents = [(0, 0, 0), (1, 2, -1), (-1, -1, -1), (4, -1, 4), (3, 3, 1), (1, 2, 3), (-1, 2, -1), (2, 2, 2)]
    counts = device.rule_counts(planes, targets, parents).cpu().numpy()
    for g, (t, p) in enumerate(zip(targets, parents)):
        assert (counts[g] == np_counts(dense, t, p)).all(), g


def test_plan_reruns_leave_the_scratch_clean(cuda):
    """The kernel zeroes its own scratch: repeated runs of one plan (no memset in between), also on
    different data, keep giving exact counts."""
    from scbonita2_b200 import device
    specs, planes_a = synth.synthetic_dataset(40, 70001, 21, flip=0.1)
    _, planes_b = synth.synthetic_dataset(40, 70001, 22, flip=0.3)
    targets, parents = _groups(specs)
    plan = device.RuleCountsPlan(40, targets, parents)
    for packed in (planes_a, planes_b, planes_a, planes_a):
        dense = synth.planes_to_dense(packed, 70001)
        counts = plan.run(device.BitPlanes.from_packed(packed, 70001)).cpu().numpy()
        for g in range(0, 40, 3):
            assert (counts[g] == np_counts(dense, targets[g], parents[g])).all(), g
    assert int(plan.scratch.view(-1).to(dtype=__import__("torch").int32).abs().sum().item()) == 0


def test_many_groups_are_sliced(cuda):
    from scbonita2_b200 import device
    rng = np.random.default_rng(4)
    dense = (rng.random((20, 2000)) < 0.3).astype(np.uint8)
    planes = device.BitPlanes.from_dense(dense)
    n_groups = 2500   # > 1024 groups per launch
    targets = rng.integers(0, 20, n_groups)
    parents = rng.integers(-1, 20, (n_groups, 3))
    counts = device.rule_counts(planes, targets, parents).cpu().numpy()
    for g in range(0, n_groups, 37):
        assert (counts[g] == np_counts(dense, targets[g], parents[g])).all(), g


def test_error_table_all_256_tables_and_direct_kernel(cuda):
    from scbonita2_b200 import device
    specs, planes_np = synth.synthetic_dataset(6, 1237, 11, flip=0.2)
    dense = synth.planes_to_dense(planes_np, 1237)
    planes = device.BitPlanes.from_packed(planes_np, 1237)
    targets, parents = _groups(specs)
    counts = device.rule_counts(planes, targets, parents)
    group = np.repeat(np.arange(6), 256)
    lut = np.tile(np.arange(256, dtype=np.uint8), 6)
    diffs = device.rule_error_table(counts, group, lut).cpu().numpy()
    direct = device.rule_error_table_direct(planes, np.asarray(targets)[group],
                                            np.asarray(parents)[group], lut).cpu().numpy()
    assert (diffs == direct).all()
    c_np = counts.cpu().numpy()
    for i in range(0, diffs.size, 5):
        assert diffs[i] == np_diff(c_np[group[i]], int(lut[i]))
    # every candidate text of every node against the oracle port (eval per cell)
    nodes = nodes_of(specs)
    for g, node in enumerate(nodes):
        node.find_all_rule_predictions()
        rules = rule_compiler.extend_with_not_variants(node)
        for r in rules[::9]:
            d, c = ref_port.calculate_error(node, r[2], dense.astype(float))
            table = rule_compiler.compile_rule(r[2])
            assert diffs[g * 256 + table] == d and c == 1237


@pytest.mark.parametrize("name", golden_names("rules_"))
def test_golden_errors_and_inference(cuda, name, tmp_path):
    """Reference (difference, count) per candidate and the full infer_ruleset() result."""
    from scbonita2_b200 import file_paths
    from scbonita2_b200.rule_determination import RuleDetermination
    logging.disable(logging.CRITICAL)
    file_paths.set_output_root(str(tmp_path / "out"))
    g = golden(name)
    dense = matrix_of(g["matrix"])
    node_dict = {s["name"]: s["index"] for s in g["specs"]}

    nodes = nodes_of(g["specs"])
    rd = RuleDetermination(None, "net", "ds", dense, nodes, node_dict)
    assert rd.num_chunks == g["num_chunks"]
    assert (rd.chunked_data_numpy == matrix_of(g["chunked"])).all()
    for node in nodes:
        node.find_all_rule_predictions()
    for node, cands, errs in zip(nodes, g["candidates"], g["errors"]):
        for logic, (d, c) in list(zip(cands, errs))[::5]:
            got = RuleDetermination.calculate_error(node, logic, rd.chunked_planes)
            assert (int(got[0]), got[1]) == (d, c)
            assert isinstance(got[0], np.int64) and isinstance(got[1], int)

    nodes = nodes_of(g["specs"])
    rd = RuleDetermination(None, "net", "ds", dense, nodes, node_dict)
    ruleset = rd.infer_ruleset()
    got = [[r[0][0], [int(i) for i in r[0][1]], r[0][2], int(r[1]), float(r[2])] for r in ruleset]
    assert got == g["ruleset"]
    for node, after in zip(nodes, g["nodes_after"]):
        assert [[k, v] for k, v in node.predecessors.items()] == after["predecessors"]
        assert [[k, bool(v)] for k, v in node.inversions.items()] == after["inversions"]
        assert len(node.node_rules) == after["n_node_rules"]
        assert (not isinstance(node.node_rules[0], list)) == after["flat_node_rules"]
        assert node.best_rule_index == after["best_rule_index"]
    rules_txt = open(os.path.join(file_paths.file_paths["rules_output"], "ds_rules",
                                  "net_ds_rules.txt")).read()
    assert rules_txt == g["rules_txt"]


def test_population_fitness_identity_and_host_path(cuda):
    """fitness[p] == sum_n E[n][choice[p][n]] (A7 pinned through A4), device and host-buffer API."""
    from scbonita2_b200 import device
    specs, planes_np = synth.synthetic_dataset(23, 5000, 21, flip=0.1)
    dense = synth.planes_to_dense(planes_np, 5000)
    planes = device.BitPlanes.from_packed(planes_np, 5000)
    targets, parents = _groups(specs)
    counts = device.rule_counts(planes, targets, parents)
    c_np = counts.cpu().numpy()
    rng = np.random.default_rng(1)
    lut = rng.integers(0, 256, (301, 23)).astype(np.uint8)
    active = (rng.random((301, 23)) < 0.8).astype(np.uint8)
    diff, diff_pg = device.population_fitness(counts, lut, None, per_node=True)
    diff, diff_pg = diff.cpu().numpy(), diff_pg.cpu().numpy()
    for p in range(0, 301, 13):
        per = [np_diff(c_np[n], int(lut[p, n])) for n in range(23)]
        assert list(diff_pg[p]) == per and diff[p] == sum(per)
    diff_a, _ = device.population_fitness(counts, lut, active)
    assert (diff_a.cpu().numpy() == (diff_pg * active).sum(axis=1)).all()

    ws = device.HostWorkspace()
    ws.load_dense(dense)
    h_counts = ws.rule_counts(targets, parents)
    assert (h_counts == c_np).all()
    h_diff, h_pg = ws.population_fitness(lut, None, per_node=True)
    assert (h_diff == diff).all() and (h_pg == diff_pg).all()
    h_tab = ws.rule_error_table(np.arange(23), lut[0])
    assert (h_tab == diff_pg[0]).all()
    ws.load_planes(planes_np, 5000)
    assert (ws.rule_counts(targets, parents) == c_np).all()
    ws.close()


def test_ga_wrapper_matches_oracle_formula(cuda):
    from scbonita2_b200 import device
    from scbonita2_b200.genetic_algorithm import PopulationFitness
    specs, planes_np = synth.synthetic_dataset(7, 90, 31, flip=0.1)
    dense = synth.planes_to_dense(planes_np, 90).astype(float)
    nodes = nodes_of(specs)
    pf = PopulationFitness(nodes, device.BitPlanes.from_packed(planes_np, 90))
    tables = pf.candidate_tables()
    logics = [[r[2] for r in n.node_rules] for n in nodes]
    rng = np.random.default_rng(2)
    choice = np.array([[rng.integers(0, len(t)) for t in tables] for _ in range(5)])
    lut = np.array([[tables[n][choice[p, n]] for n in range(7)] for p in range(5)], dtype=np.uint8)
    raw, fits = pf.fitness_calculation(lut, current_gen=4, max_gen=15)
    sel = [[[int(choice[p, n])] for n in range(7)] for p in range(5)]
    want_raw, want_fit = ref_port.population_fitness(nodes, logics, sel, dense, 4, 15)
    assert raw == want_raw and fits == want_fit


def test_counts_are_additive_over_cell_shards(cuda):
    """Size-independent property used at full size: counts(all cells) == sum of shard counts."""
    from scbonita2_b200 import device
    specs, planes_np = synth.synthetic_dataset(40, 100_000, 41, flip=0.05, relax_steps=3)
    planes = device.BitPlanes.from_packed(planes_np, 100_000)
    targets, parents = _groups(specs)
    whole = device.rule_counts(planes, targets, parents)
    n_words = planes_np.shape[1]
    parts = 0
    for lo, hi in [(0, 1000), (1000, 1001), (1001, 2500), (2500, n_words)]:
        parts = parts + device.rule_counts(planes.word_slice(lo, hi), targets, parents)
    assert (whole == parts).all()
    assert int(whole.sum()) == 40 * 100_000


@pytest.mark.parametrize("gpu_pruning", [False, True])
@pytest.mark.parametrize("name", golden_names("ruleinf_"))
def test_rule_inference_constructor_golden(cuda, name, tmp_path, gpu_pruning):
    """The whole RuleInference constructor (CSV -> sample -> binarise -> Spearman top-3 -> GPU rule
    scoring) against the reference run with the same NumPy seed; then the pickle round trip."""
    import pickle

    import networkx as nx

    import scbonita2_b200
    from scbonita2_b200 import file_paths
    from scbonita2_b200.rule_inference import RuleInference
    logging.disable(logging.CRITICAL)
    file_paths.set_output_root(str(tmp_path / "out"))
    g = golden(name)
    graph = nx.DiGraph()
    graph.add_nodes_from(g["node_names"])
    for src, dst, key, kind in g["edges"]:
        graph.add_edge(src, dst, **{key: kind})
    csv_path = tmp_path / "data.csv"
    csv_path.write_text(g["csv"])
    np.random.seed(g["seed"])
    ri = RuleInference(str(csv_path), graph, "ds", "net", ",", g["node_indices"],
                       binarize_threshold=0.001, sample_cells=True, gpu_parent_pruning=gpu_pruning)
    assert list(ri.gene_names) == g["gene_names"] and list(ri.cv_genes) == g["cv_genes"]
    assert (np.asarray(ri.binarized_matrix.todense()).astype(np.uint8) == matrix_of(g["binarized"])).all()
    assert [[int(x) for x in p] for p in ri.predecessors] == g["predecessors"]
    for node, want in zip(ri.nodes, g["nodes"]):
        assert [[k, v] for k, v in node.predecessors.items()] == want["predecessors"]
        assert [[k, bool(v)] for k, v in node.inversions.items()] == want["inversions"]
        assert [node.best_rule[0], [int(i) for i in node.best_rule[1]], node.best_rule[2]] == want["best_rule"]
        assert node.best_rule_index == want["best_rule_index"]
    got = [[r[0][0], [int(i) for i in r[0][1]], r[0][2], int(r[1]), float(r[2])] for r in ri.ruleset]
    assert got == g["ruleset"]
    # pickles carry the reference's module paths once the aliases are installed
    scbonita2_b200.install_aliases()
    blob = pickle.dumps(ri)
    assert b"rule_inference" in blob and b"node_class" in blob and b"scbonita2_b200" not in blob
    back = pickle.loads(blob)
    assert [n.best_rule for n in back.nodes] == [n.best_rule for n in ri.nodes]


def test_genetic_algorithm_seeded_and_fitness_pinned(cuda):
    """GA loop (next-1): deterministic under a seed, survivors drawn from the offspring only with the
    NSGA-II tie rule of the disassembled reference (oracle/bytecode/), and its GPU fitness equals the
    oracle formula sum(calculate_error)/sum(count) + 0.002 * #rules * gen/max_gen."""
    from scbonita2_b200 import device
    from scbonita2_b200.genetic_algorithm import GeneticAlgorithm
    specs, planes_np = synth.synthetic_dataset(8, 120, 71, flip=0.03, relax_steps=3)
    dense = synth.planes_to_dense(planes_np, 120).astype(float)
    runs = []
    for _ in range(2):
        ga = GeneticAlgorithm(nodes_of(specs), device.BitPlanes.from_packed(planes_np, 120), seed=5,
                              generations=6, population=20, parents=10, children=10)
        runs.append((ga, ga.run()))
    (ga, res), (_, res2) = runs
    assert res["rules"] == res2["rules"] and res["history"] == res2["history"]
    assert (res["best_individual"] == res2["best_individual"]).all()
    assert [h["generation"] for h in res["history"]] == list(range(7)) and len(ga.population_bits) == 10
    # selection: groups of equal fitness best first and whole; the group that crosses k gives its first,
    # its last, then its other members (crowding distance inf, inf, 0, 0, ...)
    assert GeneticAlgorithm._select([3.0, 1.0, 2.0, 1.0, 1.0, 5.0], 2) == [1, 4]
    assert GeneticAlgorithm._select([3.0, 1.0, 2.0, 1.0, 1.0, 5.0], 4) == [1, 3, 4, 2]
    assert GeneticAlgorithm._select([1, 1, 1, 1, 1], 3) == [0, 4, 1] and GeneticAlgorithm._select([2, 1], 5) == [1, 0]
    # oracle check of three final individuals: the last generation is evaluated with (gen, generations + 1)
    logics = ga.logics
    for p in (0, 4, 9):
        bits = ga.population_bits[p]
        sel = [list(np.flatnonzero(bits[ga.offsets[n]:ga.offsets[n + 1]])) for n in range(8)]
        raw, fit = ref_port.population_fitness(ga.nodes, logics, [sel], dense, ga.generations, ga.generations + 1)
        assert ga.final_raw[p] == raw[0] and ga.final_fitness[p] == fit[0][0]


def test_population_fitness_wide_counts_path(cuda):
    """Counts at or above 2^31 (multi-billion-cell totals after an all-reduce) take the int64 path."""
    from scbonita2_b200 import device
    torch = cuda
    rng = np.random.default_rng(8)
    counts_np = rng.integers(0, 1 << 40, (9, 16)).astype(np.int64)
    counts = torch.from_numpy(counts_np).cuda()
    lut = rng.integers(0, 256, (70, 9)).astype(np.uint8)
    diff, diff_pg = device.population_fitness(counts, lut, per_node=True)
    want = np.array([[np_diff(counts_np[g], int(lut[p, g])) for g in range(9)] for p in range(70)])
    assert (diff_pg.cpu().numpy() == want).all() and (diff.cpu().numpy() == want.sum(axis=1)).all()
    tab = device.rule_error_table(counts, np.arange(9), lut[3]).cpu().numpy()
    assert (tab == want[3]).all()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_peer_exchange_sums_the_shards(cuda, world):
    """scb_peer_push_counts + scb_population_fitness_peers on ONE device, `world` logical ranks wired
    to each other: the fitness of the summed shards, bit for bit, over several epochs (both
    parities of the double-buffered slots)."""
    import torch
    from scbonita2_b200 import device
    from scbonita2_b200.distributed import PeerExchange
    rng = np.random.default_rng(31 + world)
    n_groups, n_ind = 24, 70
    group = PeerExchange.local_group(n_groups, world)
    lut = torch.from_numpy(rng.integers(0, 256, (n_ind, n_groups), dtype=np.uint8)).cuda()
    try:
        for epoch in range(5):
            shards = [torch.from_numpy(rng.integers(0, 5000, (n_groups, 16))).cuda() for _ in range(world)]
            want, want_pg = device.population_fitness(sum(shards), lut, per_node=True)
            for ex, c in zip(group, shards):
                ex.push(c)
            for ex in group:
                got, got_pg = ex.population_fitness(lut, per_node=(epoch % 2 == 0))
                assert torch.equal(got, want), (world, epoch, ex.rank)
                if got_pg is not None:
                    assert torch.equal(got_pg, want_pg)
        for ex in group:
            ex.check()
    finally:
        for ex in group:
            ex.close()


def test_counts_kernel_pushes_to_peers(cuda):
    """scb_rule_counts_run_push: three cell shards on one device, each pass pushing its counts from
    the kernel's last CTA; every rank's fitness equals the unsharded one."""
    import torch
    from scbonita2_b200 import device
    from scbonita2_b200.distributed import PeerExchange, word_range
    n_nodes, n_cells, world = 30, 50_021, 3
    specs, planes_np = synth.synthetic_dataset(n_nodes, n_cells, 41, flip=0.1)
    targets, parents = _groups(specs)
    full = device.BitPlanes.from_packed(planes_np, n_cells)
    rng = np.random.default_rng(5)
    lut = torch.from_numpy(rng.integers(0, 256, (50, n_nodes), dtype=np.uint8)).cuda()
    want, _ = device.population_fitness(device.rule_counts(full, targets, parents), lut)
    group = PeerExchange.local_group(n_nodes, world)
    plans = [device.RuleCountsPlan(n_nodes, targets, parents) for _ in range(world)]
    shards = [full.word_slice(*word_range(n_cells, r, world)) for r in range(world)]
    try:
        for _ in range(3):
            for r in range(world):
                plans[r].run(shards[r], exchange=group[r])
            for r in range(world):
                got, _ = group[r].population_fitness(lut)
                assert torch.equal(got, want), r
        for ex in group:
            ex.check()
    finally:
        for ex in group:
            ex.close()


def test_spearman_top3_matches_scipy_selection(cuda):
    """GPU contingency tables + exact integer ordering pick the same three parents, in the same
    order, as the reference's scipy.stats.spearmanr loop (rule_inference.py:234-320) -- including
    duplicated rows (exact ties) and constant rows (nan -> 0)."""
    from scipy.stats import spearmanr
    from scbonita2_b200 import device, spearman
    rng = np.random.default_rng(12)
    n_rows, n_cells = 24, 3001
    dense = (rng.random((n_rows, n_cells)) < rng.uniform(0.1, 0.6, (n_rows, 1))).astype(np.uint8)
    dense[5] = dense[2]                 # exact tie between two candidates
    dense[7] = 1 - dense[2]
    dense[9] = 0                        # constant row: nan in scipy, 0 in the reference
    for r in range(10, 16):             # correlated candidates
        dense[r] = dense[0] ^ (rng.random(n_cells) < 0.05 * (r - 9))
    planes = device.BitPlanes.from_dense(dense)
    incoming = [[int(p) for p in rng.permutation(n_rows)[:rng.integers(0, 9)] if p != node]
                for node in range(n_rows)]
    incoming[0] = [2, 5, 9, 10, 11, 12, 7]
    incoming[3] = [5, 2]
    got, fallbacks = spearman.top3_parents(planes, incoming)
    assert fallbacks > 0
    for node, cands in enumerate(incoming):
        corr = []
        for p in cands:
            rho, _ = spearmanr(dense[node].astype(float).tolist(), dense[p].astype(float).tolist())
            corr.append(0 if np.isnan(rho) else rho)
        want = sorted(zip(cands, corr), reverse=True, key=lambda pair: pair[1])[:3]
        assert [p for p, _ in got[node]] == [p for p, _ in want], node
        assert [v for _, v in got[node]] == [v for _, v in want], node      # the reference's floats
    # the closed form agrees with scipy to rounding
    tabs = spearman.edge_tables(planes, [(0, 10), (0, 7), (1, 2)])
    for (a, b), t in zip([(0, 10), (0, 7), (1, 2)], tabs):
        rho, _ = spearmanr(dense[a].astype(float), dense[b].astype(float))
        assert abs(spearman.phi_from_table(*t) - rho) < 1e-12
        assert t.sum() == n_cells


# ---- the fused GA step: scb_rule_sums_run + scb_population_fitness_sums ------------------------
def _random_groups(rng, n_rows, n_groups, p_bound=0.6, targets=None):
    parents = np.full((n_groups, 3), -1, dtype=np.int32)
    for g in range(n_groups):
        bound = rng.random(3) < p_bound
        parents[g, bound] = rng.integers(0, n_rows, int(bound.sum()))
    if targets is None:
        targets = rng.integers(0, n_rows, n_groups)
    return np.asarray(targets, dtype=np.int32), parents


@pytest.mark.parametrize("n_rows,n_groups,n_cells,p_bound,seed", [
    (5, 1, 33, 1.0, 0), (12, 12, 1000, 0.6, 1), (40, 40, 4097, 0.9, 2), (64, 200, 70_001, 0.7, 3),
    (200, 200, 123_457, 0.66, 4), (300, 77, 9_999, 0.5, 5), (30, 600, 5_000, 0.4, 6), (255, 100, 31, 1.0, 7),
    (17, 17, 2_000_003, 0.8, 8)])
def test_fused_fitness_step_matches_counts_then_fitness(cuda, n_rows, n_groups, n_cells, p_bound, seed):
    """Counting kernel in sums-only mode + Moebius inversion inside the fitness kernel against the counts
    kernel + fitness kernel (itself checked against NumPy and the eval-per-cell oracle above): parent holes in any
    pattern, several groups per target row, rows nobody targets, repeated passes on one plan."""
    import torch
    from scbonita2_b200 import device
    rng = np.random.default_rng(100 + seed)
    targets, parents = _random_groups(rng, n_rows, n_groups, p_bound)
    planes_np = synth.random_planes(n_rows, n_cells, 50 + seed, density=rng.uniform(0.05, 0.9, n_rows))
    planes = device.BitPlanes.from_packed(planes_np, n_cells)
    plan = device.RuleCountsPlan(n_rows, targets, parents)
    assert plan.fused_supported
    lut = torch.from_numpy(rng.integers(0, 256, (37, n_groups), dtype=np.uint8)).cuda()
    active = (rng.random((37, n_groups)) < 0.7).astype(np.uint8)
    want, want_pg = device.population_fitness(plan.run(planes).clone(), lut, per_node=True)
    want_act, _ = device.population_fitness(plan.counts, lut, active=active)
    for rep in range(3):
        got, got_pg = plan.fitness_step(planes, lut, per_node=(rep == 1), chained=(rep == 2))
        assert torch.equal(got, want), rep
        if got_pg is not None:
            assert torch.equal(got_pg, want_pg)
    got_act, _ = plan.fitness_step(planes, lut, active=active)
    assert torch.equal(got_act, want_act)
    # the scratch is clean again: the counts path still gives the same counts on the same plan
    assert torch.equal(plan.run(planes), device.rule_counts(planes, targets, parents))


def test_fused_step_refuses_what_it_cannot_hold(cuda):
    from scbonita2_b200 import device
    rng = np.random.default_rng(3)
    targets, parents = _random_groups(rng, 50, 900)
    plan = device.RuleCountsPlan(50, targets, parents)
    assert not plan.fused_supported                   # the fitness kernel's shared-memory copy of the sums does not fit
    planes = device.BitPlanes.from_packed(synth.random_planes(50, 100, 1), 100)
    with pytest.raises(ValueError):
        plan.fitness_step(planes, np.zeros((2, 900), dtype=np.uint8))
    assert plan.run(planes).shape == (900, 16)       # the counts path takes any number of groups


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_fused_step_sums_through_the_peer_regions(cuda, world):
    """Sharded cells on ONE device (logical ranks): every rank's sums kernel pushes its subset sums from
    its last CTA, every rank's fitness kernel sums the slots and inverts: equal to the unsharded step."""
    import torch
    from scbonita2_b200 import device
    from scbonita2_b200.distributed import PeerExchange, word_range
    n_nodes, n_cells = 30, 50_021
    specs, planes_np = synth.synthetic_dataset(n_nodes, n_cells, 41 + world, flip=0.1)
    targets, parents = _groups(specs)
    full = device.BitPlanes.from_packed(planes_np, n_cells)
    rng = np.random.default_rng(5)
    lut = torch.from_numpy(rng.integers(0, 256, (50, n_nodes), dtype=np.uint8)).cuda()
    want, want_pg = device.population_fitness(device.rule_counts(full, targets, parents), lut, per_node=True)
    group = PeerExchange.local_group(n_nodes, world)
    plans = [device.RuleCountsPlan(n_nodes, targets, parents) for _ in range(world)]
    shards = [full.word_slice(*word_range(n_cells, r, world)) for r in range(world)]
    try:
        for epoch in range(4):
            for r in range(world):
                plans[r].sums_run(shards[r], exchange=group[r], chained=(epoch > 1))
            for r in range(world):
                got, got_pg = plans[r].fitness_from_sums(lut, exchange=group[r], per_node=(epoch == 0))
                assert torch.equal(got, want), (epoch, r)
                if got_pg is not None:
                    assert torch.equal(got_pg, want_pg)
        for ex in group:
            ex.check()
    finally:
        for ex in group:
            ex.close()


def test_peer_timeout_poisons_the_results(cuda):
    """A rank whose peer never delivers must not leave plausible numbers behind: the kernel gives up
    after the configured time, writes INT64_MIN everywhere and raises the region's error flag (which
    check() reports once and clears)."""
    import torch
    from scbonita2_b200 import device
    from scbonita2_b200.distributed import PeerExchange
    rng = np.random.default_rng(9)
    n_groups = 10
    group = PeerExchange.local_group(n_groups, 2)
    try:
        for ex in group:
            ex.set_timeout(50)
        counts = torch.from_numpy(rng.integers(0, 1000, (n_groups, 16))).cuda()
        lut = torch.from_numpy(rng.integers(0, 256, (5, n_groups), dtype=np.uint8)).cuda()
        group[0].push(counts)                                  # rank 1 never pushes
        got, got_pg = group[0].population_fitness(lut, per_node=True)
        assert (got == torch.iinfo(torch.int64).min).all() and (got_pg == torch.iinfo(torch.int64).min).all()
        with pytest.raises(RuntimeError):
            group[0].check()
        group[0].check()                                       # reported once
        group[1].push(counts)                                  # the late rank arrives: the next read is whole again
        want, _ = device.population_fitness(2 * counts, lut)
        got, _ = group[0].population_fitness(lut)
        assert torch.equal(got, want)
    finally:
        for ex in group:
            ex.close()


def test_counting_kernel_chained_back_to_back_keeps_every_tile(cuda):
    """``rule_counts_kernel`` claims its tiles from a counter that the last CTA to finish claiming zeroes again.  A
    launch chained behind another launch of the same plan (programmatic dependent launch: it claims tiles before
    its ``griddepcontrol.wait``) must find the counter reset: in sums-only mode the subset sums of K chained passes
    pile up in the scratch, so they must be exactly K times the sums of one pass -- a tile claimed twice or
    dropped shows in every integer."""
    import torch
    from scbonita2_b200 import device
    n, cells = 120, 700_000
    specs = synth.random_network(n, 5)
    planes = device.BitPlanes.from_packed(synth.random_planes(n, cells, 77), cells)
    targets, parents = [s.index for s in specs], [synth.spec_parent_rows(s) for s in specs]
    n_acc = n + 11 * n

    def sums_after(k_passes, chained):
        plan = device.RuleCountsPlan(n, targets, parents)
        assert plan.fused_supported
        for _ in range(k_passes):
            plan.sums_run(planes, chained=chained)
        torch.cuda.synchronize()
        return plan.scratch[:8 * n_acc].view(torch.int64).clone()

    one = sums_after(1, False)
    assert int(one.sum().item()) > 0
    for k in (2, 37):
        assert torch.equal(sums_after(k, True), one * k)
    # and through a CUDA graph, where the chained launches really run back to back
    plan = device.RuleCountsPlan(n, targets, parents)
    plan.sums_run(planes, chained=True)
    torch.cuda.synchronize()
    plan.scratch.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(25):
            plan.sums_run(planes, chained=True)
    for rep in range(1, 4):
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(plan.scratch[:8 * n_acc].view(torch.int64), one * (25 * rep))
