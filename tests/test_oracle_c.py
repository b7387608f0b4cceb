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
