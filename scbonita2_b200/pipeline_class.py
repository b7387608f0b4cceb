"""``Pipeline`` -- the rule-inference entry of the reference (``code/pipeline_class.py``), reduced to
the part that is on the accelerated path.

The reference's ``Pipeline.__init__`` parses the shell script's arguments, downloads / parses KEGG
pathways (``kegg_parser.Pathways``, out of scope: network access and GraphML handling) and then
calls ``infer_rules(pathway_num, graph, node_indices)`` once per pathway graph
(``pipeline_class.py:116-131``).  That method -- ``RuleInference`` on the GPU, then the
``{dataset}_{organism}{pathway}.ruleset.pickle`` and ``{dataset}.cells.pickle`` files that
``importance_scores.run_full_importance_score`` and the later stages read
(``pipeline_class.py:134-182``) -- is what this mirror keeps, with the same name, arguments, file
names and pickled class paths (``scbonita2_b200.install_aliases()``).  The settings the reference
reads from argparse are constructor arguments here; pathway graphs are handed in by the caller
(``rule_inference(graphs)``), e.g. the reference's own ``Pathways`` object when it is available.
"""
from __future__ import annotations

import logging
import os
import pickle

from .cell_class import CellPopulation
from .file_paths import file_paths
from .rule_inference import RuleInference


class Pipeline:
    def __init__(self, data_file, dataset_name, datafile_sep=",", binarize_threshold=0.001, organism="hsa",
                 sample_cells=True, minimum_overlap=25, gpu_parent_pruning=False, pathway_graphs=None):
        logging.basicConfig(format='%(message)s', level=logging.INFO)
        self.dataset_name = dataset_name
        self.data_file = data_file
        self.datafile_sep = datafile_sep
        self.organism = organism
        self.sample_cells = self._convert_string_to_boolean(sample_cells)
        self.binarize_threshold = float(binarize_threshold)
        self.minOverlap = minimum_overlap
        self.gpu_parent_pruning = gpu_parent_pruning
        self.node_indices = []
        for path in file_paths.values():
            os.makedirs(path, exist_ok=True)
        if pathway_graphs:
            self.rule_inference(pathway_graphs)

    def rule_inference(self, pathway_graphs, gene_list=None):
        """``pathway_graphs``: {pathway name: networkx.DiGraph} as ``Pathways.pathway_graphs`` holds them.
        ``gene_list`` (the data file's gene column, ``Pathways.gene_list``) maps graph nodes to CSV rows;
        read from the data file when not given (``pipeline_class.py:108-131``)."""
        if gene_list is None:
            with open(self.data_file) as fh:
                next(fh)
                gene_list = [line.split(self.datafile_sep, 1)[0] for line in fh if line.strip()]
        for pathway_num, graph in pathway_graphs.items():
            if len(graph.nodes()) >= self.minOverlap:
                self.node_indices = [gene_list.index(node) for node in graph.nodes()]
                self.infer_rules(pathway_num, graph, self.node_indices)
            else:
                logging.info(f'\t\t\tNot enough overlapping nodes for {pathway_num} '
                             f'(min {self.minOverlap}, overlap {len(graph.nodes())})')

    def infer_rules(self, pathway_num, graph, node_indices):
        """Rule determination for one pathway; writes the ruleset and cell-population pickles
        (``pipeline_class.py:134-182``)."""
        logging.info('\n-----RULE INFERENCE-----')
        logging.info(f'Pathway: {pathway_num}')
        logging.info(f'Num nodes: {len(node_indices)}')
        ruleset = RuleInference(self.data_file, graph, self.dataset_name, pathway_num, self.datafile_sep,
                                node_indices, self.binarize_threshold, self.sample_cells,
                                **({"gpu_parent_pruning": True} if self.gpu_parent_pruning else {}))

        logging.info('\nRule inference complete, saving ruleset pickle file')
        data_pickle_folder = f'{file_paths["pickle_files"]}/{self.dataset_name}_pickle_files/ruleset_pickle_files'
        os.makedirs(data_pickle_folder, exist_ok=True)
        if self.organism not in pathway_num:
            data_pickle_file_path = f'{data_pickle_folder}/{self.dataset_name}_{self.organism}{pathway_num}.ruleset.pickle'
        else:
            data_pickle_file_path = f'{data_pickle_folder}/{self.dataset_name}_{pathway_num}.ruleset.pickle'
        logging.info(f'\tSaving to {data_pickle_file_path.split("/")[-1]}')
        with open(data_pickle_file_path, "wb") as fh:
            pickle.dump(ruleset, fh)

        logging.info('Saving cell population pickle file')
        cell_population = CellPopulation(ruleset.cells)
        cell_pickle_dir = f'{file_paths["pickle_files"]}/{self.dataset_name}_pickle_files/cells_pickle_file'
        os.makedirs(cell_pickle_dir, exist_ok=True)
        with open(f'{cell_pickle_dir}/{self.dataset_name}.cells.pickle', "wb") as fh:
            pickle.dump(cell_population, fh)
        return ruleset

    def _convert_string_to_boolean(self, variable):
        return variable == "True" or variable is True
