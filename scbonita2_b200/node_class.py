"""``Node`` — the per-gene model object of the scBONITA2 pipeline (host side).

Mirrors the public surface of the reference ``Node`` (``code/node_class.py:5-133``): same
constructor, same attribute names (instances are pickled inside ``*.ruleset.pickle`` /
``*.network.pickle`` and read back by attribute, ``importance_scores.py:387-403``), same
candidate-template order (``node_class.py:45-107``) and the same PKN-inversion rewrite
(``node_class.py:109-133``).  The template enumerator is an independent implementation: it
builds read-once AND/OR trees directly instead of the reference's string-splicing cache, and
is pinned against the reference's output by ``tests/golden/templates.json``.
"""
from __future__ import annotations

from itertools import combinations

import numpy as np

_VARS = "ABCD"


class _Formula:
    """A read-once monotone formula: a leaf, or an AND/OR over >=2 children of the other kind."""

    __slots__ = ("op", "kids", "text")

    def __init__(self, op, kids=(), leaf=None):
        self.op = op          # None (leaf) | "and" | "or"
        self.kids = kids
        if op is None:
            self.text = leaf
        else:
            # children are ordered by their rendered text (plain ASCII order: '(' < 'A'),
            # which is what the reference's sorted(split(...)) produces (node_class.py:73-78)
            parts = sorted(self._as_child(k) for k in kids)
            self.text = f" {op} ".join(parts)

    def _as_child(self, kid):
        # an OR below an AND needs parentheses; an AND below an OR does not
        if self.op == "and" and kid.op == "or":
            return f"({kid.text})"
        return kid.text

    def top_level(self):
        """Rendering used in ``Node.possibilities``: the ROOT operator is double-spaced."""
        if self.op is None:
            return self.text
        parts = sorted(self._as_child(k) for k in self.kids)
        return f"  {self.op}  ".join(parts)


def _set_partitions(items):
    """All partitions of ``items`` into >=2 non-empty blocks (each partition once)."""
    items = list(items)
    first, rest = items[0], items[1:]
    out = []
    for k in range(len(rest) + 1):
        for together in combinations(rest, k):
            block = (first,) + together
            remaining = [x for x in rest if x not in together]
            if not remaining:
                continue
            out.append([block, tuple(remaining)])
            for sub in _set_partitions(remaining):
                out.append([block] + sub)
    return out


def _formulas(variables, memo):
    """Every read-once AND/OR formula using each variable in ``variables`` exactly once."""
    key = tuple(variables)
    if key in memo:
        return memo[key]
    if len(key) == 1:
        memo[key] = [_Formula(None, leaf=key[0])]
        return memo[key]
    found = {}
    for blocks in _set_partitions(key):
        per_block = [_formulas(b, memo) for b in blocks]

        def expand(i, chosen):
            if i == len(per_block):
                for op in ("and", "or"):
                    # alternation: children of an AND are leaves or ORs (and vice versa)
                    if all(c.op != op for c in chosen):
                        f = _Formula(op, tuple(chosen))
                        found.setdefault(f.top_level(), f)
                return
            for cand in per_block[i]:
                expand(i + 1, chosen + [cand])

        expand(0, [])
    ordered = sorted(found.values(), key=lambda f: f.top_level().replace("  ", " "))
    memo[key] = ordered
    return ordered


def rule_templates(num_parents: int) -> list:
    """Ordered candidate templates for a node with ``num_parents`` incoming nodes.

    1 / 4 / 17 / 69 templates for 1 / 2 / 3 / 4 parents, full-arity forms first, then the
    AB / AC / BC pair forms, then the single letters (``node_class.py:96-107``; the reference
    never adds pair forms involving ``D``).
    """
    memo: dict = {}

    def full(letters):
        return [f.top_level() for f in _formulas(tuple(letters), memo)]

    if num_parents <= 1:
        return ["A"]
    if num_parents == 2:
        return full("AB") + ["A", "B"]
    if num_parents == 3:
        return full("ABC") + full("AB") + full("AC") + full("BC") + ["A", "B", "C"]
    if num_parents == 4:
        return (full("ABCD") + full("ABC") + full("AB") + full("AC") + full("BC")
                + ["A", "B", "C"])
    raise IndexError("Num incoming nodes out of range")


class Node:
    def __init__(self, name, index, predecessors, inversions):
        self.name = name                  # gene name
        self.index = index                # row of this gene in the network's binarised matrix
        self.dataset_index = None         # set by RuleInference.create_nodes, unused downstream

        # {predecessor row index: bool inverted} from the prior-knowledge network
        self.inversions = inversions

        # {predecessor row index: gene name}, insertion order defines the A, B, C, D slots
        # (rule_determination.py:536,545-552).  A node without parents signals to itself.
        if len(predecessors) == 0:
            self.predecessors = {self.index: self.name}
        else:
            self.predecessors = predecessors

        self.possibilities = self.enumerate_possibilities()
        self.node_rules = None

        self.best_rule_index = None
        self.best_rule = None
        self.best_rule_error = None

        # logic string evaluated during network simulation (set from best_rule[2])
        self.calculation_function = None

        self.importance_score = 0
        self.importance_score_stdev = 0.0

        self.relative_abundance = None

    def enumerate_possibilities(self):
        return np.array(rule_templates(len(self.predecessors)))

    def find_all_rule_predictions(self):
        """Templates with the PKN inversions applied, as ``[name, [parent idx...], logic]``."""
        flips = [letter for letter, inverted in zip(_VARS[:3], self.inversions.values())
                 if inverted == True]  # noqa: E712 - dict values may be numpy bools
        parent_indices = [i for i in self.predecessors]
        node_rules = []
        for template in self.possibilities:
            logic = str(template).replace("  ", " ")
            for letter in flips:
                logic = logic.replace(letter, "not " + letter)
            node_rules.append([self.name, list(parent_indices), logic])
        self.node_rules = node_rules
        return node_rules
