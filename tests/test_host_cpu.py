"""CPU suite: rule compiler, C-ABI surface, and the no-oracle-in-product rule."""
import ctypes
import os
import re

import pytest

from scbonita2_b200 import _lib, rule_compiler
from scbonita2_b200.node_class import Node

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def all_three_parent_candidates():
    node = Node("X", 3, {0: "a", 1: "b", 2: "c"}, {0: False, 1: False, 2: False})
    node.find_all_rule_predictions()
    return [r[2] for r in rule_compiler.extend_with_not_variants(node)]


def test_candidate_counts():
    for k, n in ((1, 2), (2, 12), (3, 94)):
        node = Node("X", 9, {i: f"p{i}" for i in range(k)}, {i: False for i in range(k)})
        node.find_all_rule_predictions()
        assert len(rule_compiler.extend_with_not_variants(node)) == n


def test_94_candidates_have_94_distinct_tables():
    cands = all_three_parent_candidates()
    tables = [rule_compiler.compile_rule(c) for c in cands]
    assert len(set(tables)) == 94
    # spot values under the lop3 convention bit = (A<<2 | B<<1 | C)
    assert rule_compiler.compile_rule("A") == 0xF0
    assert rule_compiler.compile_rule("B") == 0xCC
    assert rule_compiler.compile_rule("C") == 0xAA
    assert rule_compiler.compile_rule("A and B or C") == (0xF0 & 0xCC) | 0xAA
    assert rule_compiler.compile_rule("not A and (B or not C)") == (~0xF0 & (0xCC | ~0xAA)) & 0xFF


def test_compile_rejects_bad_rules():
    with pytest.raises(rule_compiler.RuleCompileError):
        rule_compiler.compile_rule("A and D")
    with pytest.raises(rule_compiler.RuleCompileError):
        rule_compiler.compile_rule("A and and B")
    with pytest.raises(rule_compiler.RuleCompileError):
        rule_compiler.compile_for_slots("A or B", 1)


def test_lut_support():
    assert rule_compiler.lut_support(0xF0) == (True, False, False)
    assert rule_compiler.lut_support(rule_compiler.compile_rule("B or C")) == (False, True, True)


def declared_functions():
    text = open(os.path.join(ROOT, "include", "scbonita_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/ but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert _lib.lib().scb_abi_version() == 1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scbonita2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "ref_port" not in text and "ref_loader" not in text, f


def test_aliases_make_pickles_reference_compatible():
    import pickle
    import sys

    import scbonita2_b200
    from scbonita2_b200.cell_class import Cell, CellPopulation
    from scbonita2_b200.network_class import Network
    saved = {k: sys.modules.get(k) for k in ("node_class", "network_class", "cell_class")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        scbonita2_b200.install_aliases()
        net = Network("hsa04370")
        net.nodes = [Node("G0", 0, {}, {})]
        blob = pickle.dumps((net, CellPopulation([Cell(0)])))
        for path in (b"network_class", b"node_class", b"cell_class"):
            assert path in blob
        assert b"scbonita2_b200" not in blob
        back, cells = pickle.loads(blob)
        assert back.nodes[0].predecessors == {0: "G0"} and cells.cells[0].index == 0
    finally:
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def test_dtw_table_matches_the_restated_fastdtw():
    """scb_dtw_table_build is HOST code (no GPU): its entries must equal the oracle's restatement of
    compute_dtw_distance_pair's per-gene distance (attractor_analysis.py:332-343) -- squared difference
    below EUCLIDEAN_THRESHOLD, FastDTW (restated, parity unpinned) from there on."""
    import ctypes
    import numpy as np
    from oracle import ref_port
    from scbonita2_b200 import _lib
    for bits in (10, 9, 4):
        n = 1 << bits
        table = np.zeros((n, n), dtype=np.uint8)
        _lib.check(_lib.lib().scb_dtw_table_build(bits, 1, 5, table.ctypes.data), "scb_dtw_table_build")
        rng = np.random.default_rng(bits)
        n_dtw = 0
        for _ in range(1500):
            px, py = int(rng.integers(0, n)), int(rng.integers(0, n))
            x = np.array([(px >> k) & 1 for k in range(bits)])
            y = np.array([(py >> k) & 1 for k in range(bits)])
            squared = int(np.sum((x - y) ** 2))
            want = squared if squared < 5 else ref_port.fastdtw(x, y, 1, dist=lambda a, b: (a - b) ** 2)[0]
            n_dtw += squared >= 5
            assert int(table[px, py]) == want, (bits, px, py)
        assert bits < 9 or n_dtw > 500
        assert (np.diag(table) == 0).all()
    # FastDTW is not symmetric in its arguments; the product keeps the reference's argument order
    big = np.zeros((1024, 1024), dtype=np.uint8)
    _lib.check(_lib.lib().scb_dtw_table_build(10, 1, 5, big.ctypes.data), "scb_dtw_table_build")
    assert (big != big.T).any()


def test_bind_host_to_device_never_raises():
    """``device.bind_host_to_device`` (NUMA placement of pinned buffers, used by bench.py) is best effort: without
    CUDA or NVML it changes nothing and says so."""
    import os
    from scbonita2_b200 import device
    before = os.sched_getaffinity(0)
    ok = device.bind_host_to_device(0)
    assert ok in (True, False)
    if not ok:
        assert os.sched_getaffinity(0) == before
