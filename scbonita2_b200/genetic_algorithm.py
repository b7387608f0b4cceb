"""GA population fitness on the GPU (historic ``CustomDeap.fitness_calculation``; the GA survives
only as bytecode ``code/__pycache__/updated_deap_class.cpython-36.pyc`` — SURVEY fact A / A7).

An individual switches on one or more candidate rules per node.  Its raw fitness is
``sum(differences) / sum(counts)`` over the selected rules, where ``(difference, count)`` is
``calculate_error`` of that rule — the same per-node joint counts as rule scoring, so the
population costs ONE pass over the cells (``scb_rule_counts``) plus ``scb_population_fitness``.
"""
from __future__ import annotations

import numpy as np

from . import device
from .rule_compiler import compile_for_slots, extend_with_not_variants, parent_rows


class PopulationFitness:
    """Per-dataset state: node groups + their joint counts, reused across generations."""

    def __init__(self, nodes, planes: device.BitPlanes, counts_hook=None):
        self.nodes = list(nodes)
        self.planes = planes
        self.n_cells_total = planes.n_cells
        targets = [n.index for n in self.nodes]
        parents = [parent_rows(n) for n in self.nodes]
        counts = device.rule_counts(planes, targets, parents)
        self.counts = counts_hook(counts) if counts_hook is not None else counts

    def candidate_tables(self) -> list:
        """Ordered candidate truth tables per node (the GA's gene alphabet)."""
        tables = []
        for node in self.nodes:
            if node.node_rules is None or not isinstance(node.node_rules[0], list):
                node.find_all_rule_predictions()
            rules = extend_with_not_variants(node)
            n_slots = min(3, len(node.predecessors))
            tables.append(np.array([compile_for_slots(r[2], n_slots) for r in rules], dtype=np.uint8))
        return tables

    def evaluate(self, lut, active=None, per_node: bool = False):
        """Integer mismatch totals for ``lut`` [P, N] (one truth table per node per individual)."""
        return device.population_fitness(self.counts, lut, active, per_node)

    def fitness_calculation(self, lut, current_gen: int = 0, max_gen: int = 1, n_cells=None):
        """(raw_fitnesses, fitnesses) with one selected rule per node:
        raw = sum(diff) / (N * cells); fitness = raw + 0.002 * N * current_gen / max_gen."""
        diff, _ = self.evaluate(lut)
        diff = diff.cpu().numpy()
        n_rules = len(self.nodes)
        cells = int(n_cells if n_cells is not None else self.n_cells_total)
        raw = [int(d) / (n_rules * cells) for d in diff]
        fits = [(r + 0.002 * n_rules * current_gen / max_gen,) for r in raw]
        return raw, fits
