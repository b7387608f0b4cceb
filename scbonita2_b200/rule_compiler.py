"""Rule strings -> 3-input truth tables, and the ordered candidate list of a node.

The reference evaluates a candidate rule by ``eval`` of its Python text once per cell
(``rule_determination.py:554-571``) and, during simulation, by rewriting ``and/or/not`` to
``& | ~`` for numexpr (``importance_scores.py:50-57``, ``attractor_analysis.py:52-58``).  Both
are pure Boolean functions of the parent slots A, B, C, so a rule is fully described by an
8-bit truth table.  Bit ``(A<<2)|(B<<1)|C`` of the table is the rule's value for that input —
the same convention as the ``lop3.b32`` immediate, so a table can be fed straight to LOP3.

``candidate_logics`` reproduces the ORDER of the reference's candidate list
(``node_class.py:109-133`` + ``rule_determination.py:255-288,462-475``): order decides which
of several equal-error rules wins (``rule_determination.py:165``).
"""
from __future__ import annotations

from functools import lru_cache
from itertools import product

import numpy as np

SLOTS = "ABC"


class RuleCompileError(ValueError):
    pass


@lru_cache(maxsize=None)
def compile_rule(logic: str) -> int:
    """Truth table (0..255) of ``logic`` over the slots A, B, C.

    The table is computed with BOTH reference evaluators — Python ``and/or/not`` on float
    0.0/1.0 operands (rule inference) and ``& | ~`` on NumPy bools (simulation) — and the two
    must agree, otherwise one table could not stand for the rule on both halves of the path.
    """
    if "D" in logic:
        # a 4th parent only appears through the self-loop append (rule_determination.py:245-250)
        # and no rule the reference scores or simulates ever names it
        raise RuleCompileError(f"rule {logic!r} references slot D; only A, B, C are supported")
    try:
        py_code = compile(logic, "<rule>", "eval")
        bit_code = compile(logic.replace("and", "&").replace("or", "|").replace("not", "~"),
                           "<rule-bitwise>", "eval")
    except SyntaxError as exc:
        raise RuleCompileError(f"cannot parse rule {logic!r}: {exc}") from None
    table = 0
    for idx in range(8):
        a, b, c = (idx >> 2) & 1, (idx >> 1) & 1, idx & 1
        try:
            as_python = eval(py_code, {"__builtins__": {}},
                             {"A": float(a), "B": float(b), "C": float(c)})
            as_bits = eval(bit_code, {"__builtins__": {}},
                           {"A": np.bool_(a), "B": np.bool_(b), "C": np.bool_(c)})
        except Exception as exc:
            raise RuleCompileError(f"cannot evaluate rule {logic!r}: {exc}") from None
        # the reference compares the prediction with the 0/1 target row, so only ==1 matters
        py_bit = 1 if as_python == 1 else 0
        if as_python not in (0, 1):
            raise RuleCompileError(f"rule {logic!r} evaluates to non-Boolean {as_python!r}")
        if py_bit != int(bool(as_bits)):
            raise RuleCompileError(f"rule {logic!r}: python and bitwise evaluators disagree")
        table |= py_bit << idx
    return table


def lut_support(table: int) -> tuple:
    """Which of the slots (A, B, C) the truth table actually depends on."""
    bits = [(table >> i) & 1 for i in range(8)]
    dep_a = any(bits[i] != bits[i | 4] for i in range(8) if not i & 4)
    dep_b = any(bits[i] != bits[i | 2] for i in range(8) if not i & 2)
    dep_c = any(bits[i] != bits[i | 1] for i in range(8) if not i & 1)
    return dep_a, dep_b, dep_c


def not_variants(logic: str) -> list:
    """All assignments of ``not`` to the variables of ``logic`` (plain form first).

    Same ordering as ``RuleDetermination.generate_not_combinations``
    (``rule_determination.py:255-288``): variables in A, B, C, D order, ``itertools.product``
    over (var, 'not var').
    """
    present = [v for v in "ABCD" if v in logic]
    bare = logic.replace("not ", "")
    out = []
    for choice in product(*[(v, "not " + v) for v in present]):
        text = bare
        for var, repl in zip(present, choice):
            text = text.replace(var, repl)
        out.append(text)
    return out


def extend_with_not_variants(node) -> list:
    """Grow ``node.node_rules`` in place to the full candidate list and return it.

    Mirrors ``calculate_refined_errors`` (``rule_determination.py:459-475``): the variants are
    generated from the PKN-adjusted templates present BEFORE extension, and a variant is
    appended only if the exact ``[name, parents, logic]`` triple is not in the list yet.
    The in-place growth of ``node.node_rules`` is a reference side effect that later code
    (and pickles) can observe, so it is kept.
    """
    rules = node.node_rules
    parents = [p for p in node.predecessors]
    seeds = [r[2] for r in rules]
    seen = {r[2] for r in rules if r[0] == node.name and r[1] == parents}
    for seed in seeds:
        for logic in not_variants(seed):
            if logic not in seen:
                seen.add(logic)
                rules.append([node.name, list(parents), logic])
    return rules


def parent_rows(node) -> tuple:
    """Matrix rows bound to the slots A, B, C (-1 = slot unused): the first three keys of
    ``node.predecessors`` in insertion order (``rule_determination.py:536,545-552``)."""
    keys = [int(k) for k in node.predecessors][:3]
    return tuple(keys + [-1] * (3 - len(keys)))


def compile_for_slots(logic: str, n_slots: int) -> int:
    """Truth table of ``logic`` for a node that binds only ``n_slots`` parent slots.

    A rule naming an unbound slot raises ``NameError`` in the reference's ``eval``
    (``rule_determination.py:559-562``); the same condition is an error here.
    """
    for i, letter in enumerate(SLOTS):
        if letter in logic and i >= n_slots:
            raise RuleCompileError(f"rule {logic!r} uses slot {letter} but the node binds "
                                   f"only {n_slots} parent slot(s)")
    return compile_rule(logic)
