"""GPU parity: knock-out / knock-in importance sweep and per-cell trajectories."""
import logging
import random

import numpy as np
import pytest

from oracle import ref_port
from tests.helpers import golden, golden_names, matrix_of, nodes_of
from tools import synth

pytestmark = pytest.mark.gpu


def _sweep(nodes, dense, steps=50):
    from scbonita2_b200 import device
    from scbonita2_b200.importance_scores import compile_network
    for node in nodes:
        node.calculation_function = node.best_rule[2]
    parents, luts = compile_network(nodes)
    planes = device.BitPlanes.from_dense(dense)
    raw, missing, keyerr = device.ko_ki_sweep(planes, parents, luts, steps=steps)
    return raw.cpu().numpy().tolist(), missing.cpu().numpy().tolist(), int(keyerr.item())


@pytest.mark.parametrize("name", golden_names("sweep_"))
def test_sweep_golden(cuda, name, tmp_path):
    from scbonita2_b200 import file_paths
    from scbonita2_b200.importance_scores import CalculateImportanceScore
    logging.disable(logging.CRITICAL)
    file_paths.set_output_root(str(tmp_path / "out"))
    g = golden(name)
    dense = matrix_of(g["matrix"])
    nodes = nodes_of(g["specs"], with_logic=True)
    for node, logic in zip(nodes, g["logics"]):
        node.best_rule[2] = logic

    # kernel-level: counts of missing attractors per condition (also for the KeyError case)
    raw, missing, keyerr = _sweep(nodes, dense[:, g["sample"]])
    assert missing[0] == sum(1 for v in g["len_normal"] if v == 0)
    for i in range(len(nodes)):
        assert missing[1 + 2 * i] == sum(1 for v in g["len_ko"][i] if v == 0), i
        assert missing[2 + 2 * i] == sum(1 for v in g["len_ki"][i] if v == 0), i

    # API-level: same seed -> same sample -> same raw totals and normalised scores
    random.seed(g["seed"])
    import scipy.sparse as sp
    calc = CalculateImportanceScore(nodes, sp.csr_matrix(dense.astype(float)).astype(bool), "net", "ds")
    assert calc.cell_sample_indices == g["sample"]
    if g["error"] == "KeyError":
        assert keyerr > 0
        with pytest.raises(KeyError):
            calc.calculate_importance_scores()
        return
    assert keyerr == 0 and raw == g["raw"]
    scores = calc.calculate_importance_scores()
    assert [calc.raw_importance_scores[n.name] for n in nodes] == g["raw"]
    got = [float(scores[n.name]) for n in nodes]
    assert np.allclose(got, g["scores"], rtol=0, atol=1e-6)      # north_star tolerance
    assert got == g["scores"]                                     # and in fact bit-identical
    assert [float(n.importance_score) for n in nodes] == g["scores"]


def _cyclic_specs(seed, n_nodes):
    """Random network with a high share of inhibitory edges and self-free feedback: many cycles."""
    return synth.random_network(n_nodes, seed, p_inhibit=0.5, p_source=0.05)


@pytest.mark.parametrize("n_nodes,n_cells,seed,steps", [
    (3, 1, 1, 50), (6, 33, 2, 50), (12, 64, 3, 50), (17, 100, 4, 50), (25, 40, 5, 50),
    (9, 70, 6, 7), (9, 70, 7, 2), (14, 65, 8, 50), (33, 32, 9, 50)])
def test_sweep_random_vs_oracle(cuda, n_nodes, n_cells, seed, steps):
    specs = _cyclic_specs(seed, n_nodes)
    planes_np = synth.random_planes(n_nodes, n_cells, seed + 100, density=0.5)
    dense = synth.planes_to_dense(planes_np, n_cells)
    nodes = nodes_of(specs, with_logic=True)
    raw, missing, keyerr = _sweep(nodes, dense, steps)
    try:
        want_raw, want_missing = ref_port.importance_raw_counts(nodes, dense.astype(bool), steps)
    except KeyError:
        assert keyerr > 0
        return
    assert keyerr == 0
    assert missing == want_missing
    assert raw == want_raw


def test_sweep_long_cycles_and_step_window(cuda):
    """Johnson rings: period 2n.  Found iff mu + lambda <= steps - 1."""
    for n, steps in [(4, 50), (12, 50), (24, 50), (25, 50), (24, 49), (12, 25), (12, 24), (30, 64)]:
        specs = []
        for i in range(n):
            p = (i - 1) % n
            inv = i == 0
            specs.append(synth.NodeSpec(f"G{i}", i, {p: f"G{p}"}, {p: inv}, "not A" if inv else "A"))
        rng = np.random.default_rng(n)
        dense = (rng.random((n, 37)) < 0.5).astype(np.uint8)
        nodes = nodes_of(specs, with_logic=True)
        raw, missing, keyerr = _sweep(nodes, dense, steps)
        try:
            want_raw, want_missing = ref_port.importance_raw_counts(nodes, dense.astype(bool), steps)
        except KeyError:
            assert keyerr > 0, (n, steps)
            continue
        assert (raw, missing, keyerr) == (want_raw, want_missing, 0), (n, steps)


def test_sweep_is_additive_over_cell_shards_and_host_path(cuda):
    from scbonita2_b200 import device
    from scbonita2_b200.importance_scores import compile_network
    specs, planes_np = synth.synthetic_dataset(40, 5000, 77, flip=0.3, relax_steps=1, p_inhibit=0.4)
    nodes = nodes_of(specs, with_logic=True)
    parents, luts = compile_network(nodes)
    planes = device.BitPlanes.from_packed(planes_np, 5000)
    whole = device.ko_ki_sweep(planes, parents, luts)
    out = None
    n_words = planes_np.shape[1]
    for lo, hi in [(0, 7), (7, 64), (64, 65), (65, n_words)]:
        out = device.ko_ki_sweep(planes.word_slice(lo, hi), parents, luts, out=out)
    for a, b in zip(whole, out):
        assert (a == b).all()
    ws = device.HostWorkspace()
    ws.load_planes(planes_np, 5000)
    raw, missing, keyerr = ws.ko_ki_sweep(parents, luts)
    assert (raw == whole[0].cpu().numpy()).all() and (missing == whole[1].cpu().numpy()).all()
    assert keyerr[0] == int(whole[2].item())
    ws.close()


def test_sweep_wide_network_uses_32_slot_kernel(cuda):
    """> 215 nodes no longer fit 64 slots of shared memory; the 32-slot instantiation runs."""
    specs = synth.random_network(260, 5, p_inhibit=0.3)
    planes_np = synth.random_planes(260, 48, 6, density=0.5)
    dense = synth.planes_to_dense(planes_np, 48)
    nodes = nodes_of(specs, with_logic=True)
    raw, missing, keyerr = _sweep(nodes, dense, 20)
    try:
        want_raw, want_missing = ref_port.importance_raw_counts(nodes, dense.astype(bool), 20)
    except KeyError:
        assert keyerr > 0
        return
    assert (raw, missing, keyerr) == (want_raw, want_missing, 0)


@pytest.mark.parametrize("name", golden_names("traj_"))
def test_trajectories_golden(cuda, name):
    from scbonita2_b200.attractor_analysis import simulate_trajectories, vectorized_run_simulation
    g = golden(name)
    dense = matrix_of(g["matrix"])
    nodes = nodes_of(g["specs"], with_logic=True)
    for cell, want in zip(g["cells"], g["trajectories"]):
        assert vectorized_run_simulation(nodes, dense[:, cell]) == want
    batch = simulate_trajectories(nodes, dense, g["cells"])
    for m, want in enumerate(g["trajectories"]):
        assert batch[m].T.tolist() == want


def test_trajectories_many_cells_vs_oracle(cuda):
    from scbonita2_b200.attractor_analysis import simulate_trajectories
    specs = _cyclic_specs(12, 21)
    planes_np = synth.random_planes(21, 1500, 13, density=0.4)
    dense = synth.planes_to_dense(planes_np, 1500)
    nodes = nodes_of(specs, with_logic=True)
    cells = list(range(0, 1500, 7)) + [1499, 3, 3]
    got = simulate_trajectories(nodes, dense, cells)
    hist = ref_port.run_simulation(nodes, dense[:, cells].astype(bool), 20)   # (steps, nodes, cells)
    assert (got == np.transpose(hist, (2, 1, 0)).astype(np.uint8)).all()


def test_attractor_dumps_cross_check_the_sweep(cuda, tmp_path):
    """next-2: the attractor dicts the reference pickles (rebuilt from GPU trajectories + a host
    first-repeat search) must reproduce the sweep kernel's integer totals through the reference's
    own scoring recipe (align + XOR + sum)."""
    import pickle

    from scbonita2_b200 import file_paths
    from scbonita2_b200.importance_scores import CalculateImportanceScore
    file_paths.set_output_root(str(tmp_path / "out"))
    g = golden("sweep_mixed.json")       # cycle lengths 4 and 12, transients
    dense = matrix_of(g["matrix"])
    nodes = nodes_of(g["specs"], with_logic=True)
    for node, logic in zip(nodes, g["logics"]):
        node.best_rule[2] = logic
    calc = CalculateImportanceScore(nodes, dense, "net", "ds", cells=None)
    raw, missing, keyerr = calc.run_sweep()
    normal = calc.vectorized_run_simulation()
    lens = [normal[c].shape[1] if c in normal else 0 for c in range(dense.shape[1])]
    assert sorted(lens) == sorted(g["len_normal"])      # same cells, sample order differs
    totals = []
    for node in nodes:
        ko = calc.vectorized_run_simulation(knockout_node=node.name)
        ki = calc.vectorized_run_simulation(knockin_node=node.name)
        total = 0
        for cell in range(dense.shape[1]):
            if cell in ko and cell in ki:
                for pert in (ki[cell], ko[cell]):
                    a, b = calc.align_attractors(pert, normal[cell])
                    total += int(np.sum(a ^ b))
        totals.append(total)
    assert totals == raw.tolist() == g["raw"]
    calc.write_intermediates()
    with open(f"{calc.intermediate_dir}/knockout_results_{nodes[0].name}.pkl", "rb") as fh:
        back = pickle.load(fh)
    assert set(back) == set(calc.vectorized_run_simulation(knockout_node=nodes[0].name))


@pytest.mark.parametrize("n_nodes,n_cells,steps,fast,tame", [
    (100, 1500, 50, None, 0.7), (100, 1500, 50, "3", 0.7), (100, 900, 30, "50", 0.7), (128, 700, 50, None, 0.7),
    (129, 700, 50, "5", 0.7), (208, 600, 50, None, 0.7), (212, 500, 50, "4", 0.7), (64, 2100, 50, None, 0.7),
    (65, 1000, 17, None, 0.7), (90, 800, 50, None, 0.2), (150, 500, 50, "8", 0.2), (200, 400, 50, None, 0.0)])
def test_sweep_general_tables_and_parent_holes_vs_c_oracle(cuda, monkeypatch, n_nodes, n_cells, steps, fast, tame):
    """Arbitrary 8-bit truth tables and parent slots bound in any pattern (e.g. only B, or A and C):
    the kernels canonicalise the node (bound slots packed left, table rewritten) and, per node count,
    take the register ring with 4 / 8 / 13 nodes per warp or the shared-memory ring (> 208 nodes);
    every combination of passes (SCB_SWEEP_FAST_STEPS) must give the cell-by-cell oracle's integers."""
    from oracle import cport
    from scbonita2_b200 import device
    rng = np.random.default_rng(n_nodes * 1000 + steps)
    parents = np.full((n_nodes, 3), -1, dtype=np.int32)
    for n in range(n_nodes):
        bound = rng.random(3) < 0.6
        parents[n, bound] = rng.integers(0, n_nodes, int(bound.sum()))
    # `tame` = share of monotone-ish tables (the dynamics settle); the rest are arbitrary (long cycles)
    luts = np.where(rng.random(n_nodes) < tame, rng.choice([0xF0, 0xCC, 0xAA, 0x80, 0xFE, 0xE8, 0x0F, 0x33], n_nodes),
                    rng.integers(0, 256, n_nodes)).astype(np.uint8)
    planes_np = synth.random_planes(n_nodes, n_cells, 7 + n_nodes, density=0.4)
    if fast is None:
        monkeypatch.delenv("SCB_SWEEP_FAST_STEPS", raising=False)
    else:
        monkeypatch.setenv("SCB_SWEEP_FAST_STEPS", fast)
    raw, missing, keyerr = device.ko_ki_sweep(device.BitPlanes.from_packed(planes_np, n_cells), parents, luts, steps)
    w_raw, w_missing, w_key = cport.importance_raw(synth.planes_to_dense(planes_np, n_cells), parents, luts, steps)
    assert int(keyerr.item()) == w_key
    assert (missing.cpu().numpy() == w_missing).all()
    if w_key == 0:
        assert (raw.cpu().numpy() == w_raw).all()


@pytest.mark.parametrize("n_nodes,koki", [(5, "plane"), (16, "plane"), (17, "plane"), (40, "plane"), (49, "plane"),
                                          (70, "plane"), (96, "plane"), (111, "plane"), (113, "plane"), (140, "plane"),
                                          (160, "plane"), (176, "plane"), (177, "plane"), (193, "plane"), (208, "plane"),
                                          (70, "old"), (177, "old"), (200, "old"),
                                          (23, "old-compact"), (24, "old-compact"), (100, "old-compact"), (200, "old-compact"),
                                          (209, "auto"), (216, "auto")])
def test_sweep_koki_kernels_vs_c_oracle(cuda, monkeypatch, n_nodes, koki):
    """The KO/KI pass on the plane words has two implementations: ``koki_plane_kernel<CNT>`` (two CTAs per SM,
    CNT = ceil(N / 16) nodes per warp, node rows padded to 16 * CNT: every CNT from 1 to 13 is hit here, with
    and without padding rows) and ``sweep_kernel<64, MODE_KOKI>`` (``SCB_SWEEP_KOKI=old``).  Both must give
    the cell-by-cell oracle's integers on mixed dynamics (fixed points, period 2, longer cycles, no repeat)."""
    from oracle import cport
    from scbonita2_b200 import device
    rng = np.random.default_rng(n_nodes)
    n_cells, steps = 1300, 50
    parents = np.full((n_nodes, 3), -1, dtype=np.int32)
    for n in range(n_nodes):
        bound = rng.random(3) < 0.6
        parents[n, bound] = rng.integers(0, n_nodes, int(bound.sum()))
    luts = np.where(rng.random(n_nodes) < 0.6, rng.choice([0xF0, 0xCC, 0xAA, 0x80, 0xFE, 0xE8, 0x0F, 0x33], n_nodes),
                    rng.integers(0, 256, n_nodes)).astype(np.uint8)
    planes_np = synth.random_planes(n_nodes, n_cells, 11 + n_nodes, density=0.4)
    monkeypatch.delenv("SCB_SWEEP_FAST_STEPS", raising=False)
    for var in ("SCB_SWEEP_KOKI", "SCB_SWEEP_RECHECK", "SCB_SWEEP_CLEANUP"):
        monkeypatch.delenv(var, raising=False)
    if koki == "old":
        monkeypatch.setenv("SCB_SWEEP_KOKI", "old")
    if koki == "old-compact":   # sweep_kernel<64, MODE_RECHECK / MODE_CLEANUP> instead of koki_compact_kernel
        monkeypatch.setenv("SCB_SWEEP_RECHECK", "old")
        monkeypatch.setenv("SCB_SWEEP_CLEANUP", "old")
    raw, missing, keyerr = device.ko_ki_sweep(device.BitPlanes.from_packed(planes_np, n_cells), parents, luts, steps)
    w_raw, w_missing, w_key = cport.importance_raw(synth.planes_to_dense(planes_np, n_cells), parents, luts, steps)
    assert int(keyerr.item()) == w_key
    assert (missing.cpu().numpy() == w_missing).all()
    if w_key == 0:
        assert (raw.cpu().numpy() == w_raw).all()


def test_trajectories_with_parent_holes_vs_c_oracle(cuda):
    """Parent slots bound in any pattern and arbitrary tables through scb_trajectories_forced."""
    from oracle import cport
    from scbonita2_b200 import device
    rng = np.random.default_rng(5)
    n_nodes, n_cells = 45, 300
    parents = np.full((n_nodes, 3), -1, dtype=np.int32)
    for n in range(n_nodes):
        bound = rng.random(3) < 0.5
        parents[n, bound] = rng.integers(0, n_nodes, int(bound.sum()))
    luts = rng.integers(0, 256, n_nodes).astype(np.uint8)
    planes_np = synth.random_planes(n_nodes, n_cells, 21, density=0.5)
    dense = synth.planes_to_dense(planes_np, n_cells)
    cells = [0, 1, 31, 32, 33, 150, 299]
    got = device.trajectories(device.BitPlanes.from_packed(planes_np, n_cells), cells, parents, luts, 20).cpu().numpy()
    for k, c in enumerate(cells):
        want = cport.trajectory(dense[:, c], parents, luts, 20)     # (steps, nodes)
        assert (got[k] == np.asarray(want).T).all()


def test_trajectory_distances_vs_oracle(cuda):
    """next-3: compute_dtw_distances (attractor_analysis.py:347-377) from one pair-distance kernel,
    against the oracle's per-pair restatement; inputs are real simulated trajectories plus random
    ones (the latter reach the FastDTW branch often)."""
    from scbonita2_b200 import attractor_analysis as aa
    specs = _cyclic_specs(3, 30)
    planes_np = synth.random_planes(30, 400, 17, density=0.5)
    dense = synth.planes_to_dense(planes_np, 400)
    nodes = nodes_of(specs, with_logic=True)
    cells = list(range(0, 400, 9))
    traj = aa.simulate_trajectories(nodes, dense, cells)               # [M, N, 20]
    rng = np.random.default_rng(0)
    noise = (rng.random((23, 30, 20)) < 0.5).astype(np.uint8)
    for data, label in ((traj, "sim"), (noise, "rnd")):
        trajectories = {f"cell_{label}_{m}": {n.name: [int(v) for v in data[m, g]] for g, n in enumerate(nodes)}
                        for m in range(data.shape[0])}
        got = aa.compute_dtw_distances(trajectories)
        names = list(trajectories)
        assert list(got) == [(a, b) for i, a in enumerate(names) for b in names[i + 1:]]
        for (a, b), total in got.items():
            assert total == ref_port.dtw_distance_pair(trajectories[a], trajectories[b]), (a, b)
        a, b = names[0], names[-1]
        assert aa.compute_dtw_distance_pair(a, b, trajectories) == (a, b, got[(a, b)])
        matrix = aa.create_distance_matrix(got, names)
        assert (matrix == matrix.T).all() and matrix[0, len(names) - 1] == got[(a, b)]
