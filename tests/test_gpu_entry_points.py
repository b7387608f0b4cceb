"""GPU tests of the entry points the round-1 suite never touched: ``attractor_analysis.simulate_cells``
(CSV layout of the reference, attractor_analysis.py:120-137), ``distributed.sharded_rule_errors`` /
``sharded_importance`` (one device, logical shards), and exact parity of the importance sweep at the
sizes where the deferral lists, the 1024-cell batches and the cleanup pass actually fill -- against the
64-cells-per-word C oracle (golden-pinned in tests/test_oracle_c.py)."""
import logging
import os

import numpy as np
import pytest

from oracle import cport, ref_port
from tests.helpers import nodes_of
from tools import synth

pytestmark = pytest.mark.gpu


def test_simulate_cells_writes_the_reference_csv_layout(cuda, tmp_path):
    """One file per cell, one line per node: ``name,s1,...,s20`` (states after 1..20 updates), existing
    files are skipped, the cells drawn when none are named come from the unsimulated ones."""
    from scbonita2_b200 import file_paths
    from scbonita2_b200.attractor_analysis import simulate_cells
    from scbonita2_b200.network_class import Network
    logging.disable(logging.CRITICAL)
    fp = file_paths.set_output_root(str(tmp_path / "out"))
    specs, planes_np = synth.synthetic_dataset(14, 90, 23, flip=0.2, relax_steps=1, p_inhibit=0.4)
    dense = synth.planes_to_dense(planes_np, 90)
    network = Network("net")
    network.nodes = nodes_of(specs, with_logic=True)
    cells = [0, 5, 89, 31]
    assert list(simulate_cells(dense, cells, 4, network, "ds", "net")) == cells
    d = f'{fp["trajectories"]}/ds_net/text_files/cell_trajectories'
    assert sorted(os.listdir(d)) == sorted(f"cell_{c}_trajectory.csv" for c in cells)
    for c in cells:
        lines = open(f"{d}/cell_{c}_trajectory.csv").read().splitlines()
        want = np.asarray(ref_port.cell_trajectory(network.nodes, dense[:, c], 20)).T   # (nodes, steps)
        assert len(lines) == 14
        for node, line, row in zip(network.nodes, lines, want):
            fields = line.split(",")
            assert fields[0] == node.name and [int(v) for v in fields[1:]] == row.astype(int).tolist()
            assert len(fields) == 21
    # an existing file is left alone; unnamed cells are drawn from the ones without a file
    open(f"{d}/cell_5_trajectory.csv", "w").write("kept\n")
    np.random.seed(3)
    drawn = simulate_cells(dense, [], 6, network, "ds", "net")
    assert len(drawn) == 6 and not set(int(c) for c in drawn) & set(cells)
    assert open(f"{d}/cell_5_trajectory.csv").read() == "kept\n"
    assert len(os.listdir(d)) == 10


def test_sharded_entry_points_on_logical_shards(cuda):
    """``sharded_rule_errors`` and ``sharded_importance`` with rank / world given explicitly (no process
    group): the per-shard integer results add up to the unsharded ones."""
    import scipy.sparse as sp
    from scbonita2_b200 import device, distributed
    from scbonita2_b200.importance_scores import CalculateImportanceScore
    from scbonita2_b200.rule_determination import RuleDetermination
    logging.disable(logging.CRITICAL)
    specs, planes_np = synth.synthetic_dataset(16, 1300, 7, flip=0.15, relax_steps=2)
    dense = synth.planes_to_dense(planes_np, 1300)
    world = 3

    def scored(nodes, planes, **kw):      # score_candidates keys its result by id(node): list it in node order
        out = RuleDetermination.score_candidates(nodes, planes, **kw)
        return [[float(e) for e in out[id(n)]] for n in nodes]
    whole = scored(_prepared(specs), device.BitPlanes.from_dense(dense))
    parts = []
    for rank in range(world):
        collected = {}

        def hook(t, _c=collected):      # stands in for the all-reduce: keep this shard's counts
            _c["counts"] = t.clone()
            return t
        local = distributed.shard_dense(dense, rank, world)
        RuleDetermination.score_candidates(_prepared(specs), device.BitPlanes.from_dense(local), counts_hook=hook,
                                           count_override=1300)
        parts.append(collected["counts"])
    total = sum(parts)

    def summed_hook(t):
        t.copy_(total)
        return t
    nodes = _prepared(specs)
    local = distributed.shard_dense(dense, 0, world)
    again = scored(nodes, device.BitPlanes.from_dense(local), counts_hook=summed_hook, count_override=1300)
    assert again == whole
    # sharded_rule_errors itself with an explicit (rank, world): no process group -> the hook is the identity,
    # so the shard's own errors come back, normalised by the GLOBAL cell count
    own_nodes = _prepared(specs)
    own = distributed.sharded_rule_errors(own_nodes, dense, rank=1, world_size=world)
    c0, c1 = distributed.cell_range(1300, 1, world)
    direct = scored(_prepared(specs), device.BitPlanes.from_dense(dense[:, c0:c1]), count_override=1300)
    assert [[float(e) for e in own[id(n)]] for n in own_nodes] == direct
    # importance: per-shard calculators, integer totals summed by the "all-reduce"
    def calc_for(cols):
        ns = nodes_of(specs, with_logic=True)
        return CalculateImportanceScore(ns, sp.csr_matrix(dense[:, cols].astype(float)).astype(bool), "n", "d", cells=None)
    full = calc_for(slice(None))
    want = full.calculate_importance_scores()
    shard_out = []
    for rank in range(world):
        c0, c1 = distributed.cell_range(1300, rank, world)
        shard_out.append(calc_for(slice(c0, c1)).run_sweep())
    raw = sum(s[0] for s in shard_out)

    def fake_allreduce(out):
        import torch
        out[0].copy_(torch.from_numpy(raw).to(out[0].device))
        return out
    c0, c1 = distributed.cell_range(1300, 0, world)
    calc = calc_for(slice(c0, c1))
    got = calc.calculate_importance_scores(allreduce=fake_allreduce)
    assert {k: float(v) for k, v in got.items()} == {k: float(v) for k, v in want.items()}
    assert distributed.sharded_importance(calc_for(slice(None))) == want      # no process group: plain run


def _prepared(specs):
    nodes = nodes_of(specs)
    for node in nodes:
        node.find_all_rule_predictions()
    return nodes


def _compare_with_bitsliced_oracle(planes_np, n_cells, parents, luts, steps=50):
    from scbonita2_b200 import device
    raw, missing, keyerr = device.ko_ki_sweep(device.BitPlanes.from_packed(planes_np, n_cells), parents, luts, steps)
    w_raw, w_missing, w_key = cport.importance_raw_bitsliced(planes_np, n_cells, parents, luts, steps)
    assert int(keyerr.item()) == w_key
    assert (missing.cpu().numpy() == w_missing).all()
    assert (raw.cpu().numpy() == w_raw).all()


def test_sweep_exact_on_the_125k_cell_block(cuda):
    """The block of the full-size linearity test (200 nodes x 124,992 cells, imp-1M's generator): per-node
    deferral lists, 1024-cell batches of the recheck and cleanup passes and the sort by settle step all
    fill here; every integer must equal the oracle's."""
    from scbonita2_b200.importance_scores import compile_network
    block_cells = 125_000 - 8
    specs, block = synth.synthetic_dataset(200, block_cells, 5, flip=0.05, relax_steps=30)
    parents, luts = compile_network(nodes_of(specs, with_logic=True))
    _compare_with_bitsliced_oracle(block, block_cells, parents, luts)


def test_sweep_exact_on_s100k(cuda):
    """cfg2's shape: 60-node planted network x 100,000 cells, 10 % flips."""
    from scbonita2_b200.importance_scores import compile_network
    specs, planes_np = synth.synthetic_dataset(60, 100_000, 11, flip=0.10, relax_steps=30)
    parents, luts = compile_network(nodes_of(specs, with_logic=True))
    _compare_with_bitsliced_oracle(planes_np, 100_000, parents, luts)


@pytest.mark.parametrize("fast,sort", [(None, None), ("3", None), ("50", None), ("12", "0"), (None, "0")])
def test_sweep_exact_on_a_hard_k50_like_case(cuda, monkeypatch, fast, sort):
    """K50's generator (10 relaxation steps: most cells sit on cycles longer than 4 or never repeat inside
    the window) at 40,000 cells through every split of the passes: any SCB_SWEEP_FAST_STEPS, sorted or not."""
    from scbonita2_b200.importance_scores import compile_network
    for key, val in (("SCB_SWEEP_FAST_STEPS", fast), ("SCB_SWEEP_SORT", sort)):
        if val is None:
            monkeypatch.delenv(key, raising=False)
        else:
            monkeypatch.setenv(key, val)
    n_cells = 40_000
    specs = synth.random_network(120, 2)
    start = synth.random_planes(120, n_cells, 1002)
    planes_np = synth.relax_and_flip(specs, start, n_cells, 2, relax_steps=10, flip=0.05)
    parents, luts = compile_network(nodes_of(specs, with_logic=True))
    hard = cport.importance_raw_bitsliced(planes_np[:, :64], 2048, parents, luts, 50)
    assert hard[1][0] > 0 and hard[1][1:].max() > 0 and hard[2] > 0   # cells without a repeat, KeyError pairs
    _compare_with_bitsliced_oracle(planes_np, n_cells, parents, luts)


def test_every_kernel_family_once(cuda):
    """tools/sanitize_target.py -- the workload meant for compute-sanitizer (closed on this pool, see
    profiles/r2_sanitizer_closed.txt) -- as a parity run: every kernel family at small sizes, each result
    checked against NumPy / the C oracle."""
    from tools import sanitize_target
    sanitize_target.main()
