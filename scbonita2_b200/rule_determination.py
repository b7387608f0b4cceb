"""``RuleDetermination`` — rule scoring and selection, GPU-backed (reference:
``code/rule_determination.py``).

Same constructor, methods, return values and node side effects as the reference class; the
arithmetic that dominates the reference (``calculate_error``: Python ``eval`` once per rule per
column, ``rule_determination.py:509-582``) is replaced by ONE pass over the cell bitplanes on
the GPU (``scb_rule_counts``) plus a table lookup per candidate (``scb_rule_error_table``).
Selection (``:291-396,446-507``) stays on the host and works on the integer table, so every
float it compares is the same ``np.int64 / int`` quotient the reference computes.
"""
from __future__ import annotations

import logging
import os
from statistics import stdev

import numpy as np

from . import device
from .file_paths import file_paths
from .rule_compiler import compile_for_slots, extend_with_not_variants, parent_rows

ERROR_CUTOFF = 0.20          # rule_determination.py:162,331
CHUNK_ABOVE_COLUMNS = 2000   # rule_determination.py:26
CELLS_PER_CHUNK = 10         # rule_determination.py:27


def chunk_min_ones(width: int) -> int:
    """Smallest count of ones for which a chunk of ``width`` cells becomes 1.

    The reference takes ``np.mean`` of a scipy-sparse slice and tests ``>= 0.5``
    (``rule_determination.py:224-228``).  scipy scales every stored value by the float
    ``1/width`` and accumulates the scaled values, so an exact tie can land just below 0.5
    (width 12: six ones give 0.49999999999999994).  The threshold is therefore derived with
    the same float recipe instead of ``ceil(width / 2)``.
    """
    step = 1.0 / width
    acc = 0.0
    for ones in range(width + 1):
        if acc >= 0.5:
            return ones
        acc += step
    return width + 1


def chunk_layout(num_columns: int, num_chunks) -> tuple:
    """(permutation, n_chunks, chunk_size) of ``chunk_data`` (``rule_determination.py:191-209``).
    Uses the legacy global NumPy RNG exactly as the reference does (seed 42 on every call)."""
    num_chunks = int(num_chunks)
    np.random.seed(42)
    perm = np.random.permutation(num_columns)
    if num_chunks > num_columns:
        num_chunks = num_columns
    chunk_size = max(1, num_columns // num_chunks)
    return perm, num_chunks, chunk_size


class RuleDetermination:
    def __init__(self, network, network_name, dataset_name, binarized_matrix, nodes, node_dict):
        self.nodes = nodes
        self.node_dict = node_dict
        self.network = network
        self.network_name = network_name
        self.dataset_name = dataset_name
        self.binarized_matrix = binarized_matrix

        logging.basicConfig(format='%(message)s', level=logging.INFO)

        if isinstance(binarized_matrix, device.BitPlanes):
            self.planes = binarized_matrix
        else:
            self.planes = device.BitPlanes.from_dense(binarized_matrix)
        num_columns = self.planes.n_cells

        if num_columns > CHUNK_ABOVE_COLUMNS:
            self.num_chunks = round(num_columns / CELLS_PER_CHUNK)
        else:
            self.num_chunks = num_columns
        self.chunked_planes = self.chunk_data(self.num_chunks)
        self._chunked_numpy = None
        self._coarse_numpy = None

    # the reference materialises both as dense float64 arrays; here they are views on demand
    @property
    def chunked_data_numpy(self) -> np.ndarray:
        if self._chunked_numpy is None:
            self._chunked_numpy = self.chunked_planes.to_dense().astype(np.float64)
        return self._chunked_numpy

    @property
    def coarse_chunked_dataset(self) -> np.ndarray:
        # built but never read by the reference (rule_determination.py:29,33)
        if self._coarse_numpy is None:
            if self.planes.n_cells > CHUNK_ABOVE_COLUMNS:
                coarse = self.chunk_data(round(self.num_chunks / 2, 1))
            else:
                coarse = self.chunked_planes
            self._coarse_numpy = coarse.to_dense().astype(np.float64)
        return self._coarse_numpy

    def chunk_data(self, num_chunks) -> device.BitPlanes:
        """Majority-vote chunking on the GPU; returns bitplanes over the chunk columns."""
        perm, n_chunks, size = chunk_layout(self.planes.n_cells, num_chunks)
        return self.planes.chunk_majority(perm, n_chunks, size, chunk_min_ones(size))

    # ------------------------------------------------------------------------------
    def infer_ruleset(self):
        num_rows, num_columns = self.planes.n_rows, self.planes.n_cells
        chunked_rows, chunked_columns = self.chunked_planes.n_rows, self.chunked_planes.n_cells
        if chunked_columns < num_columns:
            logging.info('\n-----CHUNKING DATASET-----')
            logging.info(f"\tOriginal Data Shape: {num_rows} rows, {num_columns} columns")
            logging.info(f"\tChunked Data Shape: {chunked_rows} rows, {chunked_columns} columns")

        for node in self.nodes:
            node.find_all_rule_predictions()

        best_ruleset, ruleset_errors, node_errors = self.refine_rules()

        logging.info('\n-----RULESET-----')
        logging.info('(Genes that signal to themselves not shown)')
        self.write_ruleset(best_ruleset, ruleset_errors[0], node_errors)

        for node_index, node in enumerate(self.nodes):
            node.best_rule = best_ruleset[node_index][0]
            node.best_rule_index = best_ruleset[node_index][1]

        logging.info('Note: rules with high error likely have incoming nodes from other pathways')
        return best_ruleset

    def refine_rules(self):
        logging.info('\n-----RULE REFINEMENT-----')
        # which nodes are scored at all (rule_determination.py:145-156)
        scored = [node for node in self.nodes
                  if node.node_rules[0][1][0] != node.index and len(node.node_rules) != 0]
        errors_by_node = self.score_candidates(scored, self.chunked_planes)

        best_rules = []
        num_self_loops = 0
        for node in self.nodes:
            if id(node) not in errors_by_node:
                best_rules.append(self.handle_self_loop(node))
                num_self_loops += 1
                continue
            equivalent = self.select_best_rules(node, node.node_rules, errors_by_node[id(node)])
            if equivalent[0][2] > ERROR_CUTOFF:
                best_rules.append(self.handle_self_loop(node))
            else:
                best_rules.append(equivalent[0])

        errors = [error for _, _, error in best_rules]
        ruleset_error = [[sum(errors) / len(errors), stdev(errors), max(errors), min(errors),
                          num_self_loops]]
        return best_rules, ruleset_error, errors

    # ------------------------------------------------------------------------------
    @staticmethod
    def score_candidates(nodes, planes: device.BitPlanes, counts_hook=None,
                         count_override=None) -> dict:
        """{id(node): [difference / count per candidate]} for ``nodes`` in one GPU pass.

        Grows each node's ``node_rules`` to the full ordered candidate list first
        (``rule_determination.py:459-475``).  ``counts_hook`` (optional) receives the device
        count tensor before it is read back — the cell-sharded path all-reduces it there and
        passes the GLOBAL number of cells as ``count_override``.
        """
        if not nodes:
            return {}
        targets, parents, task_group, task_lut, spans = [], [], [], [], []
        for g, node in enumerate(nodes):
            rules = extend_with_not_variants(node)
            n_slots = min(3, len(node.predecessors))
            targets.append(node.index)
            parents.append(parent_rows(node))
            start = len(task_lut)
            for rule in rules:
                task_group.append(g)
                task_lut.append(compile_for_slots(rule[2], n_slots))
            spans.append((start, len(task_lut)))
        counts = device.rule_counts(planes, targets, parents)
        if counts_hook is not None:
            counts = counts_hook(counts)
        diffs = device.rule_error_table(counts, task_group, task_lut).cpu().numpy()
        count = int(planes.n_cells if count_override is None else count_override)
        out = {}
        for node, (lo, hi) in zip(nodes, spans):
            out[id(node)] = [np.int64(d) / count for d in diffs[lo:hi]]
        return out

    def calculate_refined_errors(self, node, chunked_dataset=None):
        """Per-node form kept for API parity (``rule_determination.py:446-507``)."""
        planes = self._as_planes(chunked_dataset) if chunked_dataset is not None else self.chunked_planes
        errors = self.score_candidates([node], planes)[id(node)]
        return self.select_best_rules(node, node.node_rules, errors)

    @classmethod
    def select_best_rules(cls, node, rules, prediction_errors):
        """PKN prioritisation -> minimum error -> most inputs (-> the reference's no-op
        'fewest and' filter, which keeps everything)."""
        rules, prediction_errors = cls.prioritize_pkn_inversion_rules(node, rules, prediction_errors)
        min_error_indices, _ = cls.find_min_error_indices(prediction_errors)
        top_rules, top_indices, top_errors = cls.maximize_incoming_connections(
            min_error_indices, rules, prediction_errors)
        return [(r, i, e) for r, i, e in zip(top_rules, top_indices, top_errors)]

    @staticmethod
    def prioritize_pkn_inversion_rules(node, rules, prediction_errors):
        """Keep only the PKN-sign-consistent rules if their mean error (+1e-5 each) is below
        20 % (``rule_determination.py:291-335``)."""
        follows = []
        for rule, error in zip(rules, prediction_errors):
            logic = rule[2]
            violates = False
            for slot, parent in zip("ABC", rule[1]):
                negated = f'not {slot}' in logic
                inverted = node.inversions[parent]
                if negated and inverted == False:  # noqa: E712 (values may be numpy bools)
                    violates = True
                if slot in logic and not negated and inverted == True:  # noqa: E712
                    violates = True
            if not violates:
                follows.append((rule, error))
        if not follows:
            # the reference logs a warning and then dies on an unbound local (:323-331)
            raise RuntimeError(f'No rules follow the inversion rules for node {node.name}')
        average_error = sum([err + 1e-5 for _, err in follows]) / len(follows)
        if average_error < ERROR_CUTOFF:
            return [r for r, _ in follows], [e for _, e in follows]
        return rules, prediction_errors

    @staticmethod
    def find_min_error_indices(prediction_errors):
        min_error = min(prediction_errors)
        return [i for i, v in enumerate(prediction_errors) if v == min_error], min_error

    @staticmethod
    def _num_inputs(logic: str) -> int:
        return sum(1 for slot in "ABC" if slot in logic)

    @classmethod
    def maximize_incoming_connections(cls, min_error_indices, rules, prediction_errors):
        sizes = [cls._num_inputs(rules[i][2]) for i in min_error_indices]
        most = max(sizes)
        keep = [i for i, n in zip(min_error_indices, sizes) if n == most and 1 <= most <= 3]
        return [rules[i] for i in keep], keep, [prediction_errors[i] for i in keep]

    @classmethod
    def minimize_incoming_connections(cls, min_error_indices, rules, prediction_errors):
        sizes = [cls._num_inputs(rules[i][2]) for i in min_error_indices]
        fewest = min(sizes)
        keep = [i for i, n in zip(min_error_indices, sizes) if n == fewest and 1 <= fewest <= 3]
        return [rules[i] for i in keep], keep, [prediction_errors[i] for i in keep]

    @staticmethod
    def handle_self_loop(node):
        """Self-signalling fallback; APPENDS the node to its own predecessor dict
        (``rule_determination.py:236-252``)."""
        node.node_rules = [node.name, [node.index], 'A']
        node.predecessors[node.index] = node.name
        node.inversions[node.index] = False
        return ([node.name, [node.index], 'A'], 0, 0)

    @staticmethod
    def generate_not_combinations(rule):
        from .rule_compiler import not_variants
        return not_variants(rule)

    # ------------------------------------------------------------------------------
    @staticmethod
    def _as_planes(dataset) -> device.BitPlanes:
        if isinstance(dataset, device.BitPlanes):
            return dataset
        return device.BitPlanes.from_dense(dataset)

    @staticmethod
    def calculate_error(node, predicted_rule, dataset):
        """(difference, count) of one rule for one node (``rule_determination.py:509-582``).

        ``dataset`` is the dense 0/1 matrix the reference passes (or a ``BitPlanes``).  A thin
        single-task wrapper over the batched GPU path, kept for call-compatibility and tests.
        """
        planes = RuleDetermination._as_planes(dataset)
        n_slots = min(4, len(node.predecessors))
        lut = compile_for_slots(predicted_rule, min(3, n_slots))
        counts = device.rule_counts(planes, [node.index], [parent_rows(node)])
        diff = device.rule_error_table(counts, [0], [lut]).cpu().numpy()[0]
        return np.int64(diff), int(planes.n_cells)

    # ------------------------------------------------------------------------------
    def write_ruleset(self, ruleset, error, node_errors):
        rule_dir = f'{file_paths["rules_output"]}/{self.dataset_name}_rules/'
        rule_path = f'{rule_dir}{self.network_name}_{self.dataset_name}_rules.txt'
        os.makedirs(rule_dir, exist_ok=True)
        name_of = {index: name for name, index in self.node_dict.items()}

        def with_gene_names(logic, gene_names):
            # two-stage replacement so that a gene name containing A/B/C is not re-substituted
            text = logic.replace('A', '{A_TEMP}').replace('B', '{B_TEMP}').replace('C', '{C_TEMP}')
            for slot, gene in zip("ABC", gene_names):
                text = text.replace('{' + slot + '_TEMP}', gene)
            return text

        with open(rule_path, 'w') as rule_file:
            for rule, _, _ in ruleset:
                gene_names = [name_of[i] for i in rule[1]]
                line = f'{rule[0]} = {with_gene_names(rule[2], gene_names)}'
                if rule[0] != with_gene_names(rule[2], gene_names):
                    logging.info(line)
                rule_file.write(line)
                rule_file.write('\n')

        logging.info('Refined Error:\n')
        logging.info(f'\tAverage = {error[0]}')
        logging.info(f'\tStdev = {error[1]}')
        logging.info(f'\tMax = {error[2]}')
        logging.info(f'\tMin = {error[3]}')
