"""Rule-recovery accuracy of an inferred rule set against a data set (BASELINE.json configs[1]:
planted Boolean network, 100k simulated cells) -- GPU-backed mirror of the reference's checker
``network_simulation/check_rules_output.py`` (``get_rules`` :50-112, ``check_logic`` :152-195,
``check_rules`` :198-231).

The reference walks every chosen cell and every rule in Python.  Here a rule's token list is compiled
ONCE into a truth table over its (at most three) distinct genes by running the checker's own
left-to-right evaluation on the 2^k input patterns -- its quirks included: no precedence, parentheses
dropped, and a pending ``not`` overwrites a pending ``and`` / ``or`` -- and the mismatches of all rules
come from one ``scb_rule_counts`` pass + ``scb_rule_error_table``.  A subset of cells is gathered on
the device (``scb_chunk_majority`` with chunks of one cell).
"""
from __future__ import annotations

import numpy as np

from . import device


def get_rules(path_or_text: str):
    """(ruleset, raw lines) as ``check_rules_output.get_rules``; accepts a path or the text itself."""
    text = path_or_text
    if "\n" not in path_or_text and "=" not in path_or_text:
        with open(path_or_text) as fh:
            text = fh.read()
    ruleset, raw = [], []
    for line in text.splitlines():
        line = line.strip()
        if not line or line.startswith('#') or '=' not in line:
            continue
        raw.append(line)
        tokens = line.split('=', 1)[1].strip().replace('(', ' ( ').replace(')', ' ) ').split()
        processed = []
        for tok in tokens:
            low = tok.lower()
            if low in ('and', 'or', 'not'):
                processed.append(low)
            elif low not in ('(', ')'):
                idx = low[4:] if low.startswith('gene') else low
                if not idx.isdigit():
                    raise ValueError(f"Invalid gene token: {tok!r}")
                processed.append(idx)
        if processed:
            ruleset.append(processed)
    return ruleset, raw


def _evaluate(tokens, value_of) -> int:
    """The checker's evaluation order (``check_logic``): left to right, one pending operator."""
    result = None
    operator = None
    for token in tokens:
        if token in ('and', 'or', 'not'):
            operator = token
            continue
        value = bool(value_of[int(token)])
        if operator == 'not':
            value = not value
            operator = None
        if result is None:
            result = value
        elif operator == 'and':
            result = result and value
        elif operator == 'or':
            result = result or value
        else:
            result = value
        operator = None
    return 1 if result else 0


def check_logic(rule, state) -> int:
    for token in rule:
        if token not in ('and', 'or', 'not') and not 0 <= int(token) < len(state):
            raise IndexError(f"Rule references gene index {token}, but state length = {len(state)}")
    return _evaluate(rule, state)


def compile_rule(tokens):
    """(parent rows [3], truth table) of a token list: slots A, B, C = its distinct genes in order of
    first appearance; table bit (A<<2 | B<<1 | C) = the checker's value on that pattern."""
    genes = []
    for token in tokens:
        if token not in ('and', 'or', 'not') and int(token) not in genes:
            genes.append(int(token))
    if len(genes) > 3:
        raise ValueError(f"rule over {len(genes)} genes: the counting kernels take at most three parents")
    rows = genes + [-1] * (3 - len(genes))
    lut = 0
    for idx in range(8):       # (the value never depends on an unbound slot: the tokens do not name it)
        bits = ((idx >> 2) & 1, (idx >> 1) & 1, idx & 1)
        if _evaluate(tokens, {g: b for g, b in zip(genes, bits)}):
            lut |= 1 << idx
    return rows, lut


def check_rules(dataset, rules, raw_rule_text=None, cell_indices=None):
    """(mismatch_count, total_checks, error_pct, num_rules) of ``rules`` (rule i predicts gene i) over the
    chosen cells of ``dataset`` (a 0/1 matrix or ``device.BitPlanes``), as ``check_rules_output.check_rules``;
    ``check_rules.per_rule`` of the last call holds the mismatches of every rule."""
    planes = dataset if isinstance(dataset, device.BitPlanes) else device.BitPlanes.from_dense(np.asarray(dataset))
    if len(rules) > planes.n_rows:
        raise IndexError("more rules than genes")
    if cell_indices is not None:
        idx = np.asarray(cell_indices, dtype=np.int64).reshape(-1)
        if idx.size and (idx.min() < 0 or idx.max() >= planes.n_cells):
            raise IndexError("cell index out of range")
        planes = planes.chunk_majority(idx, int(idx.size), 1, 1) if idx.size else None
    n_cells = planes.n_cells if planes is not None else 0
    if not rules or n_cells == 0:
        check_rules.per_rule = [0] * len(rules)
        return 0, 0, 0.0, len(rules)
    compiled = [compile_rule(r) for r in rules]
    targets = list(range(len(rules)))
    counts = device.rule_counts(planes, targets, [rows for rows, _ in compiled])
    diffs = device.rule_error_table(counts, targets, [lut for _, lut in compiled]).cpu().numpy()
    check_rules.per_rule = [int(d) for d in diffs]
    mismatches = int(diffs.sum())
    total = n_cells * len(rules)
    return mismatches, total, 100.0 * mismatches / total, len(rules)


check_rules.per_rule = []
