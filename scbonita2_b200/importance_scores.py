"""``CalculateImportanceScore`` / ``run_full_importance_score`` — knock-out / knock-in
importance scoring, GPU-backed (reference: ``code/importance_scores.py``).

The reference simulates the Boolean network 2N+1 times in a process pool, pickles per-cell
attractors to disk and scores them in Python (``importance_scores.py:47-143,195-340``).  Here
the whole sweep — synchronous updates, first-repeat attractor per cell, alignment and XOR
counting — is one call into ``scb_ko_ki_sweep``; the host only samples the cells, compiles
each node's rule into a truth table and normalises the integer totals by their maximum.
"""
from __future__ import annotations

import logging
import os
import pickle
import random

import numpy as np

from . import device
from .file_paths import file_paths
from .network_class import Network
from .rule_compiler import compile_for_slots, parent_rows


def compile_network(nodes) -> tuple:
    """(parent rows int32 [N,3], truth tables uint8 [N]) from ``node.calculation_function``.

    Slot binding follows the simulation code (``importance_scores.py:77-97``): A, B, C read the
    rows named by the first three keys of ``node.predecessors``.  The reference indexes the
    data by ``predecessors`` keys at step 0 and the node LIST by the same keys afterwards, which
    only agrees when ``nodes[i].index == i``; that is required here.
    """
    n = len(nodes)
    parents = np.full((n, 3), -1, dtype=np.int32)
    luts = np.zeros(n, dtype=np.uint8)
    for pos, node in enumerate(nodes):
        if node.index != pos:
            raise ValueError(f"node {node.name!r}: index {node.index} != list position {pos}; the "
                             "simulation assumes nodes are listed in matrix-row order")
        rows = parent_rows(node)
        bound = sum(1 for r in rows if r >= 0)
        for r in rows:
            if r >= n:
                raise IndexError(f"node {node.name!r} reads row {r} of a {n}-node network")
        parents[pos] = rows
        luts[pos] = compile_for_slots(node.calculation_function, bound)
    return parents, luts


class CalculateImportanceScore:
    def __init__(self, nodes, binarized_matrix, network_name, dataset_name, steps: int = 50,
                 cells: int | None = 500):
        self.binarized_matrix = binarized_matrix
        self.nodes = nodes
        self.STEPS = steps
        self.CELLS = cells
        self.network_name = network_name
        self.dataset_name = dataset_name

        self.intermediate_dir = (f'{file_paths["importance_score_output"]}/{self.dataset_name}'
                                 f'/intermediate_files/{self.network_name}')
        os.makedirs(self.intermediate_dir, exist_ok=True)

        if isinstance(binarized_matrix, device.BitPlanes):
            total = binarized_matrix.n_cells
        else:
            total = binarized_matrix.shape[1]
        n_cols = total if cells is None else min(cells, total)
        if cells is None:
            # cap lifted: every cell is simulated; order is irrelevant to the sums
            self.cell_sample_indices = list(range(total))
            self.planes = (binarized_matrix if isinstance(binarized_matrix, device.BitPlanes)
                           else device.BitPlanes.from_dense(binarized_matrix))
        else:
            # the reference's (unseeded) Python-random column sample (importance_scores.py:30-34)
            self.cell_sample_indices = random.sample(range(total), k=n_cols)
            if isinstance(binarized_matrix, device.BitPlanes):
                dense = binarized_matrix.to_dense()[:, self.cell_sample_indices]
            elif hasattr(binarized_matrix, "toarray"):
                dense = binarized_matrix[:, self.cell_sample_indices].toarray()
            else:
                dense = np.asarray(binarized_matrix)[:, self.cell_sample_indices]
            self.planes = device.BitPlanes.from_dense(dense)
        self.raw_importance_scores = None
        self.missing_attractors = None

    def prepare_nodes(self):
        """``node.calculation_function`` / ``incoming_node_indices`` exactly as
        ``perform_knockouts_knockins`` sets them (``importance_scores.py:212-216``)."""
        for node in self.nodes:
            incoming_nodes = node.best_rule[1][:]
            node.calculation_function = node.best_rule[2]
            node.incoming_node_indices = [index for index, name in node.predecessors.items()
                                          if name in incoming_nodes]

    def run_sweep(self, allreduce=None):
        """Integer results of the whole sweep as NumPy int64: (raw[N], missing[2N+1], keyerror).
        ``allreduce`` (optional) is applied to the device tensors first (cell-sharded runs)."""
        self.prepare_nodes()
        parents, luts = compile_network(self.nodes)
        out = device.ko_ki_sweep(self.planes, parents, luts, steps=self.STEPS)
        if allreduce is not None:
            out = allreduce(out)
        raw, missing, keyerror = (t.cpu().numpy() for t in out)
        return raw, missing, int(keyerror[0])

    def calculate_importance_scores(self, allreduce=None):
        logging.info('\n-----RUNNING NETWORK SIMULATION-----')
        raw, missing, keyerror = self.run_sweep(allreduce)
        self.missing_attractors = missing
        if keyerror:
            # a cell repeats under KO and KI but not in the unperturbed run: the reference
            # indexes normal_signaling[cell] there and dies (importance_scores.py:281)
            raise KeyError(f'{keyerror} (cell, node) pairs have knock-out and knock-in attractors '
                           'but no normal attractor; increase the step count')
        logging.info('\n-----CALCULATING IMPORTANCE SCORES-----')
        for pos, node in enumerate(self.nodes):
            lost = int(missing[1 + 2 * pos]) + int(missing[2 + 2 * pos])
            if lost:
                logging.error(f'\nERROR: No knockout attractor found for {lost} cell runs in node '
                              f'{node.name}. Increase step size\n')
        self.raw_importance_scores = {node.name: int(v) for node, v in zip(self.nodes, raw)}
        max_importance_score = max(self.raw_importance_scores.values())
        if max_importance_score == 0:
            raise ZeroDivisionError('division by zero')  # importance_scores.py:316-317
        scaled = {name: np.float64(score) / np.float64(max_importance_score)
                  for name, score in self.raw_importance_scores.items()}
        for node in self.nodes:
            node.importance_score = scaled[node.name]
        return scaled

    # ---- attractor dumps (reference intermediates; off the hot path) ------------------------
    def vectorized_run_simulation(self, knockout_node=None, knockin_node=None):
        """{cell position: bool array (nodes, lambda+1)} — the attractors the reference pickles as
        ``knockout_results_{node}.pkl`` / ``knockin_results_{node}.pkl`` / ``normal_signaling.pkl``
        (``importance_scores.py:47-132``).  The simulation runs on the GPU
        (``scb_trajectories_forced``); the first-repeat search over the returned 50-step histories
        is plain NumPy on the host.  ``calculate_importance_scores`` never needs these dumps (the
        sweep kernel scores in place); they exist for consumers of the intermediate files and as an
        independent cross-check of the sweep."""
        if any(node.calculation_function is None for node in self.nodes):
            self.prepare_nodes()
        parents, luts = compile_network(self.nodes)
        names = [node.name for node in self.nodes]
        forced, value = -1, 0
        if knockout_node in names:      # the reference tests the knock-out name first (:66-70)
            forced, value = names.index(knockout_node), 0
        elif knockin_node in names:
            forced, value = names.index(knockin_node), 1
        cells = np.arange(self.planes.n_cells)
        hist = device.trajectories(self.planes, cells, parents, luts, steps=self.STEPS,
                                   forced_node=forced, forced_value=value).cpu().numpy()   # [cells, N, steps]
        attractors = {}
        for cell in range(hist.shape[0]):
            states = hist[cell]
            for i in range(1, states.shape[1]):
                same = np.flatnonzero((states[:, :i] == states[:, i:i + 1]).all(axis=0))
                if same.size:
                    attractors[cell] = states[:, same[0]:i + 1].astype(bool)
                    break
        return attractors

    def write_intermediates(self):
        """Write ``normal_signaling.pkl`` and the per-node knock-out / knock-in pickles into
        ``intermediate_dir`` exactly where the reference keeps them (:219-246)."""
        self.save_to_disk(self.vectorized_run_simulation(), f'{self.intermediate_dir}/normal_signaling.pkl')
        for node in self.nodes:
            self.save_to_disk(self.vectorized_run_simulation(knockout_node=node.name),
                              f'{self.intermediate_dir}/knockout_results_{node.name}.pkl')
            self.save_to_disk(self.vectorized_run_simulation(knockin_node=node.name),
                              f'{self.intermediate_dir}/knockin_results_{node.name}.pkl')

    @staticmethod
    def align_attractors(ko_attractor, ki_attractor):
        """Trim two (nodes, steps) arrays to their common extent (``importance_scores.py:325-340``)."""
        rows = min(ko_attractor.shape[0], ki_attractor.shape[0])
        cols = min(ko_attractor.shape[1], ki_attractor.shape[1])
        return ko_attractor[:rows, :cols], ki_attractor[:rows, :cols]

    @staticmethod
    def save_to_disk(data, filename):
        with open(filename, 'wb') as f:
            pickle.dump(data, f)

    @staticmethod
    def load_from_disk(filename):
        with open(filename, 'rb') as f:
            return pickle.load(f)


def run_full_importance_score(dataset_name, network_names, steps: int = 50, cells: int | None = 500):
    """Drop-in for ``importance_scores.run_full_importance_score`` (``:343-427``): reads
    ``*.ruleset.pickle``, writes ``{network}_importance_score.txt`` and ``*.network.pickle``.
    The graph figure of the reference is only drawn when matplotlib is importable."""
    ruleset_dir = f'{file_paths["pickle_files"]}/{dataset_name}_pickle_files/ruleset_pickle_files/'
    seen = []
    for file_name in os.listdir(ruleset_dir):
        network_name = file_name.split(dataset_name + "_")[1].split(".ruleset.pickle")[0]
        print(f'Network name: {network_name}')
        seen.append(network_name)
        if network_name not in network_names:
            logging.debug(f'Skipping {network_name}')
            continue
        out_root = f'{file_paths["importance_score_output"]}/{dataset_name}'
        text_dir, png_dir, svg_dir = (f'{out_root}/text_files', f'{out_root}/png_files',
                                      f'{out_root}/svg_files')
        for d in (text_dir, png_dir, svg_dir):
            os.makedirs(d, exist_ok=True)
        score_path = f'{text_dir}/{network_name}_importance_score.txt'
        if os.path.exists(score_path):
            logging.info(f'Importance scores for {network_name} already exist, using cached files')
            continue

        logging.info(f'\nLoading: {dataset_name}_{network_name}.ruleset.pickle')
        with open(f'{ruleset_dir}/{file_name}', "rb") as fh:
            ruleset = pickle.load(fh)
        network = Network(name=f'{network_name}')
        network.nodes = ruleset.nodes
        network.rulesets = ruleset.ruleset
        network.network = ruleset.graph
        network.dataset = ruleset.binarized_matrix

        logging.info(f'Calculating importance score for network {network_name}')
        calculator = CalculateImportanceScore(network.nodes, network.dataset.astype(bool),
                                              network_name, dataset_name, steps=steps, cells=cells)
        calculator.calculate_importance_scores()

        logging.info(f'Saving importance scores to file: {network_name}_importance_score.txt')
        with open(score_path, 'w') as fh:
            fh.write("gene\timportance_score\n")
            fh.write("\n".join(f"{node.name}\t{round(node.importance_score, 3)}"
                               for node in network.nodes))
        try:
            import matplotlib.pyplot as plt
            fig = ruleset.plot_graph_from_graphml(network.network)
            fig.savefig(f'{png_dir}/{file_name}.png', bbox_inches='tight', format='png')
            fig.savefig(f'{svg_dir}/{file_name}.png', bbox_inches='tight', format='svg')
            plt.close(fig)
        except Exception as exc:  # plotting is outside the hot path and optional here
            logging.info(f'Skipping importance score figure ({type(exc).__name__})')

        network_folder = f'{file_paths["pickle_files"]}/{dataset_name}_pickle_files/network_pickle_files'
        os.makedirs(network_folder, exist_ok=True)
        network_file_path = f'{network_folder}/{dataset_name}_{network_name}.network.pickle'
        with open(network_file_path, 'wb') as fh:
            pickle.dump(network, fh)
        logging.info(f'\tSaved to {network_file_path}')

    if not [item for item in seen if item in network_names]:
        print(f'Network names: {network_names}')
        raise Exception('ERROR: Pathways specific do not exist in the ruleset.pickle folder. '
                        'Check spelling and try again')
