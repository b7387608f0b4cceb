"""CPU restatement of the reference's rule-recovery checker.  TEST INFRASTRUCTURE ONLY.

Follows ``network_simulation/check_rules_output.py``:

* :func:`parse_rules`  -- ``get_rules`` (:50-112): one ``GeneX = GeneA AND GeneB OR NOT GeneC`` line per
  gene, right-hand side tokenised on whitespace, parentheses DROPPED, ``GeneN`` -> ``"N"``
* :func:`check_logic`  -- ``check_logic`` (:152-195): tokens evaluated left to right WITHOUT precedence;
  ``not`` applies to the next gene token only, and a pending ``not`` replaces a pending ``and`` / ``or``
  (so ``A and not B`` evaluates to ``not B``: the operator variable is overwritten) -- kept as it is
* :func:`check_rules`  -- ``check_rules`` (:198-231): mismatches of every rule against its own gene's
  value over the chosen cells; error_pct = 100 * mismatches / (cells x rules)

Pinned against the reference by ``tests/golden/rulecheck_*.json`` (``oracle/make_golden.py`` imports the
unmodified module and records its outputs).
"""
from __future__ import annotations

import numpy as np


def parse_rules(text: str):
    """(ruleset, raw lines): ruleset[i] is the token list of the i-th kept line."""
    ruleset, raw = [], []
    for line in text.splitlines():
        line = line.strip()
        if not line or line.startswith('#') or '=' not in line:
            continue
        _, right = line.split('=', 1)
        raw.append(line)
        tokens = right.strip().replace('(', ' ( ').replace(')', ' ) ').split()
        processed = []
        for tok in tokens:
            low = tok.lower()
            if low in ('and', 'or', 'not'):
                processed.append(low)
            elif low in ('(', ')'):
                continue
            else:
                idx = low[4:] if low.startswith('gene') else low
                if not idx.isdigit():
                    raise ValueError(f"Invalid gene token: {tok!r}")
                processed.append(idx)
        if processed:
            ruleset.append(processed)
    return ruleset, raw


def check_logic(rule, state) -> int:
    result = None
    operator = None
    for token in rule:
        if token in ('and', 'or', 'not'):
            operator = token
            continue
        idx = int(token)
        if idx < 0 or idx >= len(state):
            raise IndexError(f"Rule references gene index {idx}, but state length = {len(state)}")
        value = bool(state[idx])
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
    if result is None:
        return 0
    return 1 if result else 0


def check_rules(dataset: np.ndarray, rules, cell_indices):
    """(mismatch_count, total_checks, error_pct, num_rules) and the per-rule mismatch list."""
    per_rule = [0] * len(rules)
    total = 0
    for cell in cell_indices:
        state = dataset[:, cell]
        for gene, rule in enumerate(rules):
            if check_logic(rule, state) != int(state[gene]):
                per_rule[gene] += 1
            total += 1
    mism = sum(per_rule)
    return mism, total, (100.0 * mism / total if total else 0.0), len(rules), per_rule
