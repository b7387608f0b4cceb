"""``RuleInference`` — data ingest + node construction in front of the GPU rule scoring
(reference: ``code/rule_inference.py:17-417``).

Host-side mirror: CSV rows of the network's genes -> (optionally sampled) cells -> csr matrix ->
binarisation -> top-3 predecessors per node by Spearman correlation -> ``Node`` objects ->
``RuleDetermination(...).infer_ruleset()`` (which runs on the GPU).  Attribute names are kept
because instances are pickled as ``{dataset}_{network}.ruleset.pickle`` and read back by
attribute (``pipeline_class.py:160-170``, ``importance_scores.py:386-403``).  Only the arithmetic
in ``RuleDetermination`` is on the accelerated path; everything here runs once per network.
"""
from __future__ import annotations

import copy
import csv
import logging

import networkx as nx
import numpy as np
import scipy.sparse as sparse
from scipy.stats import spearmanr

from .cell_class import Cell
from .node_class import Node
from .rule_determination import RuleDetermination


class RuleInference:
    """Class for single-cell experiments"""

    def __init__(self, data_file, graph, dataset_name, network_name, sep, node_indices,
                 binarize_threshold=0.001, sample_cells=True, gpu_parent_pruning=False):
        # gpu_parent_pruning (not in the reference): take the Spearman top-3 of every node from ONE
        # pass of scb_rule_counts over the bitplanes (spearman.py) instead of one scipy call per edge;
        # same parents in the same order, and the reference's floats for the stored correlations
        self.gpu_parent_pruning = bool(gpu_parent_pruning)
        self.node_indices = node_indices
        self.dataset_name = dataset_name
        self.network_name = network_name
        self.binarize_threshold = binarize_threshold
        self.max_samples = 15000
        self.cells = []

        self.predecessors_final = []
        self.rvalues = []
        self.predecessors = []
        self.num_successors = []
        self.graph = graph

        self.node_list = list(graph.nodes)
        self.node_dict = {self.node_list[i]: i for i in range(len(self.node_list))}
        self.rule_graph = nx.empty_graph(0, create_using=nx.DiGraph)

        logging.info('\n-----EXTRACTING AND FORMATTING DATA-----')
        logging.info(f'Extracting cell expression data from "{data_file}"')
        self.cell_names, self.gene_names, self.data = self._extract_data(data_file, sep, sample_cells,
                                                                         node_indices)
        self.sparse_matrix = sparse.csr_matrix(self.data)
        logging.info('\tCreated sparse matrix')
        self.sparse_matrix.eliminate_zeros()

        if self.data.shape[0] > 0:
            logging.info('\tBinarized sparse matrix')
            self.binarized_matrix = self._binarize(self.sparse_matrix, binarize_threshold)
        else:
            raise ValueError("No samples selected for binarization")

        self.filterData(self.binarize_threshold)
        self.calculate_node_information()
        self.nodes, self.deap_individual_length = self.create_nodes()
        self.create_cells()  # the reference discards the result as well

        rule_determination = RuleDetermination(self.graph, self.network_name, self.dataset_name,
                                               self.binarized_matrix, self.nodes, self.node_dict)
        self.ruleset = rule_determination.infer_ruleset()

    @staticmethod
    def _binarize(matrix, threshold):
        """``sklearn.preprocessing.binarize(csr, threshold, copy=True)``: stored values above the
        threshold become 1.0, everything else is dropped (``rule_inference.py:72``)."""
        if threshold < 0:
            raise ValueError("Cannot binarize a sparse matrix with threshold < 0")
        out = matrix.copy().tocsr()
        keep = out.data > threshold
        out.data[keep] = 1
        out.data[~keep] = 0
        out.eliminate_zeros()
        return out

    def _extract_data(self, data_file, sep, sample_cells, node_indices):
        """Rows whose (0-based, header excluded) position is in ``node_indices``, in FILE order;
        cells sampled with the global NumPy RNG whenever ``sample_cells`` or >= 15000 cells
        (``rule_inference.py:103-161``; with ``sample_cells=True`` a smaller data set is merely
        shuffled)."""
        with open(data_file, 'r') as file:
            reader = csv.reader(file, delimiter=sep)
            cell_names = next(reader)[1:]
            cell_count = len(cell_names)
            if cell_count >= self.max_samples or sample_cells:
                logging.info(f'\tRandomly sampling {self.max_samples} cells...')
                sampled_cell_indices = np.random.choice(range(cell_count), replace=False,
                                                        size=min(self.max_samples, cell_count))
                logging.info(f'\t\tNumber of cells: {len(sampled_cell_indices)}')
            else:
                sampled_cell_indices = range(cell_count)
                logging.info(f'\tLoading all {len(sampled_cell_indices)} cells...')

            wanted = set(node_indices)
            data = np.empty((len(node_indices), len(sampled_cell_indices)), dtype="float")
            gene_names = []
            filled = 0
            columns = [int(c) + 1 for c in sampled_cell_indices]   # +1 skips the gene-name column
            for i, row in enumerate(reader):
                if i in wanted:
                    gene_names.append(row[0])
                    data[filled, :] = [float(row[c]) for c in columns]
                    filled += 1
            logging.info(f'\tNumber of genes: {len(gene_names)}')
            logging.info(f'\tNumber of cells: {len(cell_names)}')
            return cell_names, gene_names, data

    def filterData(self, threshold):
        """Genes whose coefficient of variation reaches the threshold (``rule_inference.py:163-175``)."""
        self.cv_genes = []
        if threshold is not None:
            for i in range(0, self.sparse_matrix.get_shape()[0]):
                row = list(self.sparse_matrix.getrow(i).todense())
                if np.std(row) / np.mean(row) >= threshold:
                    self.cv_genes.append(self.gene_names[i])
        else:
            self.cv_genes = copy.deepcopy(self.gene_names)

    def plot_graph_from_graphml(self, network):
        import matplotlib.pyplot as plt   # optional: plotting is outside the hot path

        values = [node.importance_score for node in self.nodes]
        lo, hi = min(values), max(values)
        scaled = [((v - lo) / (hi - lo)) * (0.7 - 0.1) + 0.1 for v in values]
        node_colors = [plt.cm.Greys(v) for v in scaled]
        pos = nx.spring_layout(network, k=1)
        fig, ax = plt.subplots(figsize=(12, 8))
        nx.draw_networkx_nodes(network, pos, node_color=node_colors, node_size=500, ax=ax)
        nx.draw_networkx_edges(network, pos, edge_color='gray', ax=ax)
        nx.draw_networkx_labels(network, pos, font_color="black", font_size=10, ax=ax)
        ax.set_title("Importance Score for Each Node in the Network")
        ax.set_axis_off()
        return fig

    # ---- predecessors ------------------------------------------------------------------
    def calculate_node_information(self):
        self._gpu_top3 = None
        if self.gpu_parent_pruning:
            from . import device, spearman
            planes = device.BitPlanes.from_dense(self.binarized_matrix)
            incoming = [[self.node_list.index(p) for p in self.graph.predecessors(name)]
                        for name in self.node_list]
            self._gpu_top3, _ = spearman.top3_parents(planes, incoming)
        for node_num, _ in enumerate(self.node_list):
            chosen = self.find_predecessors(self.rule_graph, self.node_list, self.graph, node_num)
            self.predecessors.append([self.node_list.index(name) for name, _ in chosen])
        self._gpu_top3 = None

    def calculate_spearman_correlation(self, node, predecessors_temp):
        """Spearman correlation of the node's binarised row with each candidate parent's row; rows
        are addressed by GRAPH position (``rule_inference.py:234-260``)."""
        node_row = self.binarized_matrix[node, :].todense().tolist()[0]
        correlations = []
        for parent in predecessors_temp:
            parent_row = self.binarized_matrix[self.node_list.index(parent), :].todense().tolist()[0]
            rho, _ = spearmanr(node_row, parent_row)
            correlations.append(0 if np.isnan(rho) else rho)
        return correlations

    def find_predecessors(self, rule_graph, node_list, graph, node_index):
        """Top three incoming nodes by correlation, descending, ties in graph order
        (``rule_inference.py:263-320``)."""
        target = node_list[node_index]
        incoming = list(graph.predecessors(target))
        if getattr(self, "_gpu_top3", None) is not None:
            chosen = [(node_list[row], value) for row, value in self._gpu_top3[node_index]]
            top_values = sorted((value for _, value in chosen), reverse=True)
        else:
            correlations = self.calculate_spearman_correlation(node_index, incoming)
            chosen = sorted(zip(incoming, correlations), reverse=True, key=lambda pair: pair[1])[:3]
            top_values = sorted(correlations, reverse=True)[:3]

        self.num_successors.append(len(list(graph.successors(target))))
        self.rvalues.append(top_values)
        self.predecessors_final.append([name for name, _ in chosen])

        for parent, weight in chosen:
            attrs = graph[parent][target]
            for key in ("interaction", "signal"):
                if key in attrs:
                    rule_graph.add_edge(parent, target, weight=weight, activity=attrs[key])
        return chosen

    def calculate_inversion_rules(self, node_predecessors, node_index):
        """{parent index: True if the PKN edge is inhibitory ('i')} (``rule_inference.py:323-364``)."""
        inversion_rules = {}
        target = self.node_list[node_index]
        for incoming_node in list(node_predecessors):
            attrs = self.graph[self.node_list[incoming_node]][target]
            if "interaction" in attrs:
                inversion_rules[incoming_node] = attrs["interaction"] == "i"
            elif "signal" in attrs:
                inversion_rules[incoming_node] = attrs["signal"] == "i"
            else:
                # multigraph-style attribute dicts: the last inner value decides, as in the reference
                for _, inner in attrs.items():
                    for _attribute, value in inner.items():
                        inversion_rules[incoming_node] = value == "i"
        return inversion_rules

    def create_nodes(self):
        gene_row = {gene_name: i for i, gene_name in enumerate(self.gene_names)}
        name_of = {index: name for name, index in self.node_dict.items()}
        nodes = []
        for node_index, node_name in enumerate(self.node_list):
            parent_indices = self.predecessors[node_index] if node_index < len(self.predecessors) else []
            predecessors = {index: name_of[index] for index in parent_indices}
            inversions = self.calculate_inversion_rules(predecessors, node_index)
            node = Node(node_name, node_index, predecessors, inversions)
            node.dataset_index = gene_row.get(node_name)
            nodes.append(node)
        return nodes, 0

    def create_cells(self):
        full_matrix = np.asarray(self.binarized_matrix.todense())
        cells = []
        for cell_index, cell_name in enumerate(self.cell_names):
            cell = Cell(cell_index)
            cell.name = cell_name
            if cell_index < full_matrix.shape[1]:
                for row_num in range(full_matrix.shape[0]):
                    cell.expression[self.gene_names[row_num]] = full_matrix[row_num, cell_index]
            cells.append(cell)
        return cells
