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


class GeneticAlgorithm:
    """Seeded GA over multi-rule individuals, fitness on the GPU.

    The GA only survives in the reference as Python-3.6 bytecode
    (``code/__pycache__/updated_deap_class.cpython-36.pyc``); its disassembly is committed as
    ``oracle/bytecode/updated_deap_class_py36.dis.txt`` (made by ``oracle/pyc36_dis.py``) and is what this
    class follows -- source line numbers below are the ones recorded in that bytecode:

    * generation loop (``genetic_algorithm``, 83-165): the start population is evaluated with
      ``(current_gen, max_gen) = (0, generations)``; then for ``gen = 1 .. generations``: ``offspring =
      __varOrAdaptive(population, ..., gen / generations, ...)``, every offspring is evaluated with
      ``(gen, generations + 1)``, the last generation's offspring are the result, otherwise
      ``population[:] = select(offspring, parent_population_size)`` -- survivors come from the OFFSPRING
      only, there is no elitism;
    * variation (``__varOrAdaptive``, 949-975): per child one uniform draw: ``< cx`` -> two-point
      crossover on node boundaries of two sampled parents, the first child kept (``__cxTwoPointNode``,
      978-1020); ``< cx + mut`` -> adaptive bit-flip mutation of a clone (``__mutFlipBitAdapt``);
      otherwise a parent is copied;
    * fitness (``fitness_calculation``, 892-946): ``sum(difference) / sum(count)`` over every selected
      rule of every node, plus ``0.002 * (#selected rules) * current_gen / max_gen``; minimised
      (``FitnessMin``, ONE weight of -1.0: ``make_toolbox`` 648-649 -- the per-node weight tuple built
      just above it is never used);
    * selection (``__selNSGA2`` 727-756 with ``__sortNondominatedAdapt`` / ``__dominated`` 834-854):
      "dominates" compares ``mean(wvalues)``, i.e. the single fitness value, so the non-dominated fronts
      are the groups of individuals with EQUAL fitness in ascending order: NSGA-II selection degenerates
      to truncation by fitness.  What remains of it is the tie rule: fronts are taken whole, and the
      front that crosses ``k`` is ordered by crowding distance -- with identical fitness values its
      first and last member get ``inf`` and the rest ``0.0`` (``__assignCrowdingDist`` 857-890), so the
      stable descending sort yields first, last, then the others in order.  ``_select`` does exactly that;
    * start individuals (``__genBits``, 686-717): random bits, per node at most five and at least one
      set bit.

    The mutation operator is restated in simplified form (the node to rewrite is drawn with probability
    proportional to its error, ``__selectMutNode`` 1118-1137; the reference additionally biases towards
    nodes with many upstream candidates), and the reference's draws come from ``random`` / ``np.random``
    streams that a NumPy ``default_rng`` cannot replay.  PARITY UNPINNED against the historic code (no
    runnable source); the fitness values are pinned to ``calculate_error`` through the oracle identity,
    and a run is a pure function of ``seed``.  Recovered constants: 15 generations, start
    population 50, 25 parents, 25 children, crossover 0.1, mutation 0.9, bit-flip 0.5.
    """

    def __init__(self, nodes, planes: device.BitPlanes, seed: int = 0, generations: int = 15,
                 population: int = 50, parents: int = 25, children: int = 25, cx: float = 0.1,
                 mut: float = 0.9, bit_flip: float = 0.5, max_rules_per_node: int = 5,
                 counts_hook=None, n_cells=None):
        self.nodes = list(nodes)
        self.rng = np.random.default_rng(seed)
        self.generations, self.population = generations, population
        self.n_parents, self.n_children = parents, children
        self.cx, self.mut, self.bit_flip = cx, mut, bit_flip
        self.max_rules = max_rules_per_node
        self.scorer = PopulationFitness(self.nodes, planes, counts_hook)
        self.n_cells = int(n_cells if n_cells is not None else planes.n_cells)
        tables = self.scorer.candidate_tables()
        self.logics = [[r[2] for r in node.node_rules] for node in self.nodes]
        self.offsets = np.concatenate([[0], np.cumsum([len(t) for t in tables])]).astype(np.int64)
        task_group = np.repeat(np.arange(len(tables)), [len(t) for t in tables])
        task_lut = np.concatenate(tables)
        # one table of integer mismatches for every candidate rule of every node: the whole data
        # set is read once, every generation afterwards only touches this table
        self.task_diff = device.rule_error_table(self.scorer.counts, task_group, task_lut)
        self.task_error = self.task_diff.cpu().numpy() / self.n_cells
        self.history = []

    # ---- individuals -------------------------------------------------------------------
    def _repair(self, bits):
        for n in range(len(self.nodes)):
            seg = bits[self.offsets[n]:self.offsets[n + 1]]
            on = np.flatnonzero(seg)
            if on.size == 0:
                seg[self.rng.integers(0, seg.size)] = 1
            elif on.size > self.max_rules:
                seg[self.rng.choice(on, on.size - self.max_rules, replace=False)] = 0
        return bits

    def _random_individual(self):
        bits = self.rng.integers(0, 2, size=int(self.offsets[-1])).astype(np.uint8)
        return self._repair(bits)

    def evaluate(self, pop, generation, max_gen=None):
        diff, nsel = device.bitstring_fitness(self.task_diff, np.ascontiguousarray(pop))
        diff, nsel = diff.cpu().numpy(), nsel.cpu().numpy()
        raw = diff / (nsel.astype(np.int64) * self.n_cells)
        max_gen = self.generations if max_gen is None else max_gen
        return raw, raw + 0.002 * nsel * generation / max_gen

    @staticmethod
    def _select(fitness, k):
        """Indices of the ``k`` survivors (``__selNSGA2`` on a single objective, see the class notes):
        groups of equal fitness, best first and whole; the group that crosses ``k`` contributes its first,
        its last, then its other members in order."""
        order = {}
        for i, f in enumerate(fitness):
            order.setdefault(float(f), []).append(i)
        chosen = []
        for value in sorted(order):
            front = order[value]
            if len(chosen) + len(front) <= k:
                chosen.extend(front)
                if len(chosen) == k:
                    break
                continue
            ranked = [front[0], front[-1]] + front[1:-1] if len(front) > 1 else front
            chosen.extend(ranked[:k - len(chosen)])
            break
        return chosen

    def _crossover(self, a, b):
        cut = np.sort(self.rng.choice(len(self.nodes) + 1, 2, replace=False))
        lo, hi = self.offsets[cut[0]], self.offsets[cut[1]]
        child = a.copy()
        child[lo:hi] = b[lo:hi]          # two-point crossover on node boundaries
        return child

    def _mutate(self, bits):
        # nodes whose selected rules fit worst are the most likely to be rewritten
        err = np.array([self.task_error[self.offsets[n]:self.offsets[n + 1]][
            bits[self.offsets[n]:self.offsets[n + 1]] == 1].mean() for n in range(len(self.nodes))])
        weight = err + 1e-9
        node = self.rng.choice(len(self.nodes), p=weight / weight.sum())
        seg = bits[self.offsets[node]:self.offsets[node + 1]]
        flip = self.rng.random(seg.size) < self.bit_flip
        seg[flip] ^= 1
        return self._repair(bits)

    def run(self):
        population = np.stack([self._random_individual() for _ in range(self.population)])
        raw, fit = self.evaluate(population, 0, self.generations)
        self.history.append({"generation": 0, "best": float(fit.min()), "mean": float(fit.mean()),
                             "best_raw": float(raw[int(np.argmin(fit))])})
        for gen in range(1, self.generations + 1):
            kids = []
            for _ in range(self.n_children):            # __varOrAdaptive
                op_choice = self.rng.random()
                if op_choice < self.cx:
                    i, j = self.rng.choice(len(population), 2, replace=False)
                    child = self._crossover(population[i], population[j])
                elif op_choice < self.cx + self.mut:
                    child = self._mutate(population[self.rng.integers(0, len(population))].copy())
                else:
                    child = population[self.rng.integers(0, len(population))].copy()
                kids.append(child)
            offspring = np.stack(kids)
            raw, fit = self.evaluate(offspring, gen, self.generations + 1)
            self.history.append({"generation": gen, "best": float(fit.min()), "mean": float(fit.mean()),
                                 "best_raw": float(raw[int(np.argmin(fit))])})
            if gen == self.generations:
                population = offspring
                break
            population = offspring[self._select(fit, self.n_parents)]
        best = population[int(np.lexsort((np.arange(len(fit)), fit))[0])]
        self.population_bits = population
        self.final_raw, self.final_fitness = raw, fit
        rules = []
        for n, node in enumerate(self.nodes):
            seg = np.flatnonzero(best[self.offsets[n]:self.offsets[n + 1]])
            pick = seg[np.argmin(self.task_error[self.offsets[n] + seg])]   # first minimum
            rules.append((node.name, self.logics[n][int(pick)], float(self.task_error[self.offsets[n] + pick])))
        return {"best_individual": best, "rules": rules, "history": self.history}
