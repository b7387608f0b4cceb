"""cfg2 (rule-recovery accuracy): the restatement of the reference's checker against goldens made by
the unmodified ``network_simulation/check_rules_output.py``, and the GPU checker against both."""
import numpy as np
import pytest

from oracle import rule_check_port
from tests.helpers import golden, golden_names, matrix_of


@pytest.mark.parametrize("name", golden_names("rulecheck_"))
def test_checker_port_matches_reference_golden(name):
    g = golden(name)
    rules, raw = rule_check_port.parse_rules(g["rules_text"])
    assert rules == g["tokens"] and len(raw) == len(rules)
    dense = matrix_of(g["matrix"])
    for c, want in enumerate(g["check_logic_first_cells"]):
        assert [rule_check_port.check_logic(r, dense[:, c]) for r in rules] == want
    for res in g["results"]:
        m, t, e, k, _ = rule_check_port.check_rules(dense, rules, res["cells"])
        assert (m, t, k) == (res["mismatches"], res["total"], res["num_rules"])
        assert e == res["error_pct"]


def test_checker_quirks_are_kept():
    """No precedence, and a pending ``not`` swallows the pending ``and`` / ``or`` (check_logic :175-190)."""
    state = [1, 0, 1]
    assert rule_check_port.check_logic(["0", "and", "not", "1"], state) == 1     # = not Gene1
    assert rule_check_port.check_logic(["0", "and", "not", "2"], state) == 0     # = not Gene2, not (1 and 0)
    assert rule_check_port.check_logic(["1", "or", "0", "and", "1"], state) == 0  # (Gene1 or Gene0) and Gene1
    assert rule_check_port.check_logic([], state) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names("rulecheck_"))
def test_gpu_checker_matches_reference_golden(cuda, name):
    from scbonita2_b200 import rule_recovery
    g = golden(name)
    rules, raw = rule_recovery.get_rules(g["rules_text"])
    assert rules == g["tokens"]
    dense = matrix_of(g["matrix"])
    for c, want in enumerate(g["check_logic_first_cells"]):
        assert [rule_recovery.check_logic(r, dense[:, c]) for r in rules] == want
    for res in g["results"]:
        m, t, e, k = rule_recovery.check_rules(dense, rules, raw, res["cells"])
        assert (m, t, k) == (res["mismatches"], res["total"], res["num_rules"])
        assert e == res["error_pct"]


@pytest.mark.gpu
def test_gpu_checker_vs_port_on_random_rules(cuda):
    """Random token lists (every operator pattern, repeated genes, leading operators) on 20k cells."""
    from scbonita2_b200 import rule_recovery
    rng = np.random.default_rng(4)
    n_genes, n_cells = 25, 20_000
    dense = (rng.random((n_genes, n_cells)) < rng.uniform(0.1, 0.9, (n_genes, 1))).astype(np.uint8)
    rules = []
    for _ in range(n_genes):
        genes = rng.choice(n_genes, int(rng.integers(1, 4)), replace=False)
        tokens = []
        for i in range(int(rng.integers(1, 6))):
            if i and rng.random() < 0.9:
                tokens.append(str(rng.choice(["and", "or"])))
            if rng.random() < 0.35:
                tokens.append("not")
            tokens.append(str(int(rng.choice(genes))))
        rules.append(tokens)
    cells = sorted(int(c) for c in rng.choice(n_cells, 3_000, replace=False))
    want = rule_check_port.check_rules(dense, rules, cells)
    got = rule_recovery.check_rules(dense, rules, None, cells)
    assert got == want[:4] and rule_recovery.check_rules.per_rule == want[4]
    with pytest.raises(ValueError):
        rule_recovery.compile_rule(["0", "and", "1", "or", "2", "and", "3"])
