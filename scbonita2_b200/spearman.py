"""Top-3 parent pruning on the GPU.

Reference: ``RuleInference.find_predecessors`` / ``calculate_spearman_correlation``
(``rule_inference.py:234-320``) calls ``scipy.stats.spearmanr`` on two dense 0/1 rows for every
incoming edge of every node and keeps the three largest correlations (descending, stable).

For 0/1 rows the rank transform is affine, so Spearman's rho IS the phi coefficient of the 2x2
contingency table

    rho = (n11*n00 - n10*n01) / sqrt(n1. * n0. * n.1 * n.0)      (nan if a row is constant -> 0)

and the table is what ``scb_rule_counts`` already produces for a group (target = node, parent A =
candidate).  One pass over the cell bitplanes therefore serves every edge of the network.  The
ORDER of the candidates is decided exactly, on integers (cross-multiplied squares), so it does not
depend on floating-point noise; only when two candidates of a node are mathematically tied in a way
that matters (inside the top three, or across the cut) the reference's own float arithmetic decides
-- those few rows are unpacked and handed to ``scipy.stats.spearmanr``, as are, optionally, the
values stored for the chosen parents (``rvalues``, edge weights), so the pickled outputs carry the
reference's floats.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np


def phi_from_table(n11: int, n10: int, n01: int, n00: int) -> float:
    """Spearman's rho of two 0/1 vectors from their contingency table (nan -> 0 as the reference)."""
    r1, r0 = n11 + n10, n01 + n00          # first vector: ones / zeros
    c1, c0 = n11 + n01, n10 + n00          # second vector
    den = r1 * r0 * c1 * c0
    if den == 0:
        return 0.0
    return (n11 * n00 - n10 * n01) / math.sqrt(den)


def _exact_key(table) -> Fraction:
    """Monotone, exact image of rho: sign(num) * num^2 / den (0 for a constant row)."""
    n11, n10, n01, n00 = (int(v) for v in table)
    den = (n11 + n10) * (n01 + n00) * (n11 + n01) * (n10 + n00)
    if den == 0:
        return Fraction(0)
    num = n11 * n00 - n10 * n01
    return Fraction(num * abs(num), den)


def edge_tables(planes, edges):
    """int64 [E, 4] tables (n11, n10, n01, n00) with the TARGET as first vector, for E (target row,
    parent row) edges, from one ``scb_rule_counts`` pass."""
    from . import device
    edges = np.asarray(edges, dtype=np.int32).reshape(-1, 2)
    if edges.shape[0] == 0:
        return np.zeros((0, 4), dtype=np.int64)
    parents = np.full((edges.shape[0], 3), -1, dtype=np.int32)
    parents[:, 0] = edges[:, 1]
    counts = device.rule_counts(planes, edges[:, 0], parents).cpu().numpy()   # [E][minterm A,B,C][target]
    # parent A = 1 is minterm 4, A = 0 is minterm 0; column = target bit
    n11, n01 = counts[:, 2 * 4 + 1], counts[:, 2 * 4 + 0]      # target 1 / target 0 where parent = 1
    n10, n00 = counts[:, 1], counts[:, 0]                      # target 1 / target 0 where parent = 0
    return np.stack([n11, n10, n01, n00], axis=1).astype(np.int64)


def top3_parents(planes, incoming, reference_values: bool = True):
    """``incoming[node] = [candidate parent rows]`` (graph order) -> per node the chosen (parent,
    correlation) pairs exactly as ``sorted(zip(...), reverse=True, key=corr)[:3]`` would give them
    with scipy's correlations, and the number of edges that needed scipy to break a tie."""
    from scipy.stats import spearmanr

    edges = [(node, parent) for node, cands in enumerate(incoming) for parent in cands]
    tables = edge_tables(planes, edges)
    dense_rows = {}

    def row(r):
        if r not in dense_rows:
            sub = planes.words[r:r + 1]
            from .device import BitPlanes
            dense_rows[r] = BitPlanes(sub, 1, planes.n_cells).to_dense()[0].astype(float).tolist()
        return dense_rows[r]

    def scipy_rho(node, parent):
        rho, _ = spearmanr(row(node), row(parent))
        return 0 if np.isnan(rho) else rho

    chosen, fallbacks, at = [], 0, 0
    for node, cands in enumerate(incoming):
        k = len(cands)
        tabs = tables[at:at + k]
        at += k
        keys = [_exact_key(t) for t in tabs]
        phis = [phi_from_table(*t) for t in tabs]
        order = sorted(range(k), key=lambda i: keys[i], reverse=True)      # stable, like the reference
        top = order[:3]
        # ties, exact or closer than double precision can tell apart (1e-9 is generous), that touch
        # the top three: the reference's own float arithmetic decides those
        tied = set()
        for a, b in zip(order, order[1:]):
            if abs(phis[a] - phis[b]) <= 1e-9 and (a in top or b in top):
                tied.update((a, b))
        if tied:
            grew = True
            while grew:       # close the class under "within 1e-9 of a member"
                grew = False
                for i in range(k):
                    if i not in tied and any(abs(phis[i] - phis[j]) <= 1e-9 for j in tied):
                        tied.add(i)
                        grew = True
            fallbacks += len(tied)
            value = {i: (scipy_rho(node, cands[i]) if i in tied else phis[i]) for i in range(k)}
            # (reverse=True keeps graph order between equal floats, as in the reference)
            order = sorted(range(k), key=lambda i: value[i], reverse=True)
            top = order[:3]
        pairs = []
        for i in top:
            if reference_values:
                value = scipy_rho(node, cands[i])
            else:
                value = phi_from_table(*tabs[i])
            pairs.append((cands[i], value))
        chosen.append(pairs)
    return chosen, fallbacks
