"""CPU suite: the C restatement agrees with the Python port (itself pinned to the reference
goldens) and with the golden sweep fixtures directly."""
import numpy as np
import pytest

from oracle import cport, ref_port
from scbonita2_b200 import rule_compiler
from scbonita2_b200.importance_scores import compile_network
from tests.helpers import golden, golden_names, matrix_of, nodes_of
from tools import synth


@pytest.mark.parametrize("name", golden_names("sweep_"))
def test_c_sweep_matches_golden(name):
    g = golden(name)
    dense = matrix_of(g["matrix"])[:, g["sample"]]
    nodes = nodes_of(g["specs"], with_logic=True)
    for node, logic in zip(nodes, g["logics"]):
        node.calculation_function = logic
    parents, luts = compile_network(nodes)
    raw, missing, keyerr = cport.importance_raw(dense, parents, luts, 50)
    assert missing[0] == sum(1 for v in g["len_normal"] if v == 0)
    if g["error"] == "KeyError":
        assert keyerr > 0
        return
    assert keyerr == 0 and raw.tolist() == g["raw"]


def test_c_rule_difference_matches_port():
    specs, planes = synth.synthetic_dataset(9, 143, 4, flip=0.2)
    dense = synth.planes_to_dense(planes, 143)
    nodes = nodes_of(specs)
    for node in nodes:
        node.find_all_rule_predictions()
        for rule in rule_compiler.extend_with_not_variants(node)[::11]:
            want, _ = ref_port.calculate_error(node, rule[2], dense.astype(float))
            got = cport.rule_difference(dense, node.index, rule_compiler.parent_rows(node),
                                        rule_compiler.compile_rule(rule[2]))
            assert got == want


def test_c_sweep_and_trajectory_match_port_on_cyclic_network():
    specs = synth.random_network(11, 5, p_inhibit=0.5, p_source=0.05)
    planes = synth.random_planes(11, 40, 6, density=0.5)
    dense = synth.planes_to_dense(planes, 40)
    nodes = nodes_of(specs, with_logic=True)
    parents, luts = compile_network(nodes)
    raw, missing, keyerr = cport.importance_raw(dense, parents, luts, 50)
    try:
        want_raw, want_missing = ref_port.importance_raw_counts(nodes, dense.astype(bool), 50)
        assert keyerr == 0 and raw.tolist() == want_raw and missing.tolist() == want_missing
    except KeyError:
        assert keyerr > 0
    traj = cport.trajectory(dense[:, 3], parents, luts, 20)
    assert traj.tolist() == ref_port.cell_trajectory(nodes, dense[:, 3], 20)


# ---- the 64-cells-per-word restatement (oracle/c/scb_oracle_bs.c): pinned to the same goldens ----
@pytest.mark.parametrize("name", golden_names("sweep_"))
def test_bitsliced_c_sweep_matches_golden(name):
    g = golden(name)
    dense = matrix_of(g["matrix"])[:, g["sample"]]
    nodes = nodes_of(g["specs"], with_logic=True)
    for node, logic in zip(nodes, g["logics"]):
        node.calculation_function = logic
    parents, luts = compile_network(nodes)
    planes = synth.dense_to_planes(dense)
    for threads in (1, 3):
        raw, missing, keyerr = cport.importance_raw_bitsliced(planes, dense.shape[1], parents, luts, 50, threads=threads)
        assert missing[0] == sum(1 for v in g["len_normal"] if v == 0)
        for i in range(len(nodes)):
            assert missing[1 + 2 * i] == sum(1 for v in g["len_ko"][i] if v == 0)
            assert missing[2 + 2 * i] == sum(1 for v in g["len_ki"][i] if v == 0)
        if g["error"] == "KeyError":
            assert keyerr > 0
            continue
        assert keyerr == 0 and raw.tolist() == g["raw"]


@pytest.mark.parametrize("n_nodes,n_cells,steps,seed", [(7, 1, 50, 0), (11, 63, 50, 1), (11, 64, 50, 2), (23, 65, 50, 3),
                                                        (40, 777, 50, 4), (16, 200, 7, 5), (16, 200, 2, 6), (30, 130, 64, 7)])
def test_bitsliced_c_sweep_matches_cell_by_cell(n_nodes, n_cells, steps, seed):
    """Arbitrary truth tables and parent holes (long cycles, cells without a repeat, KeyError cells):
    the word-parallel restatement and the cell-by-cell one must give the same integers."""
    rng = np.random.default_rng(seed)
    parents = np.full((n_nodes, 3), -1, dtype=np.int32)
    for n in range(n_nodes):
        bound = rng.random(3) < 0.6
        parents[n, bound] = rng.integers(0, n_nodes, int(bound.sum()))
    luts = np.where(rng.random(n_nodes) < 0.6, rng.choice([0xF0, 0xCC, 0xAA, 0x80, 0xFE, 0xE8, 0x0F], n_nodes),
                    rng.integers(0, 256, n_nodes)).astype(np.uint8)
    planes = synth.random_planes(n_nodes, n_cells, seed + 50, density=0.4)
    a = cport.importance_raw(synth.planes_to_dense(planes, n_cells), parents, luts, steps)
    b = cport.importance_raw_bitsliced(planes, n_cells, parents, luts, steps, threads=2)
    assert a[2] == b[2] and (a[1] == b[1]).all() and (a[0] == b[0]).all()
