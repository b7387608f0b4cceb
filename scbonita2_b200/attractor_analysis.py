"""Per-cell trajectory simulation, GPU-backed (reference: ``code/attractor_analysis.py:43-244``).

The simulation half of the reference module is on the hot path; of what follows it, the pairwise
trajectory distance (``compute_dtw_distance_pair`` / ``compute_dtw_distances``,
``attractor_analysis.py:304-377``) is mirrored too; clustering and reporting are
out of scope.  ``vectorized_run_simulation`` keeps the
reference signature (one cell column in, ``list[steps][nodes]`` of ints out) and
``simulate_cells`` batches every requested cell into ONE ``scb_trajectories`` call instead of a
process pool with one Python simulation per cell, then writes the same
``cell_{i}_trajectory.csv`` files (``attractor_analysis.py:132-137``).
"""
from __future__ import annotations

import logging
import os

import numpy as np

from . import device
from .file_paths import file_paths
from .importance_scores import compile_network

STEPS = 20  # attractor_analysis.py:44


def vectorized_run_simulation(nodes: list, cell_column: list, steps: int = STEPS):
    """``steps`` synchronous updates from one cell's expression column."""
    column = (np.asarray(cell_column).reshape(-1) != 0).astype(np.uint8)
    if column.size != len(nodes):
        raise ValueError("cell_column must hold one value per node")
    planes = device.BitPlanes.from_dense(column.reshape(-1, 1))
    parents, luts = compile_network(nodes)
    traj = device.trajectories(planes, [0], parents, luts, steps=steps)   # [1, N, steps]
    return traj[0].cpu().numpy().T.astype(int).tolist()


def simulate_trajectories(nodes, dataset, cell_indices, steps: int = STEPS) -> np.ndarray:
    """uint8 ``[len(cell_indices), nodes, steps]`` for many cells at once."""
    planes = dataset if isinstance(dataset, device.BitPlanes) else device.BitPlanes.from_dense(dataset)
    parents, luts = compile_network(nodes)
    return device.trajectories(planes, cell_indices, parents, luts, steps=steps).cpu().numpy()


def simulate_cells(dataset_array, cells_to_simulate, num_simulations: int, network,
                   dataset_name: str, network_name: str):
    """Simulate ``num_simulations`` cells and write their trajectory CSVs; returns the cells."""
    logging.info(f'\tSimulating {num_simulations} cell trajectories')
    trajectory_dir = (f'{file_paths["trajectories"]}/{dataset_name}_{network_name}'
                      '/text_files/cell_trajectories')
    os.makedirs(trajectory_dir, exist_ok=True)
    existing = set()
    for filename in os.listdir(trajectory_dir):
        if filename.startswith("cell_") and filename.endswith("_trajectory.csv"):
            try:
                existing.add(int(filename.split('_')[1]))
            except (IndexError, ValueError):
                logging.warning(f"Skipping file {filename}: could not extract cell index.")
    n_total = dataset_array.n_cells if isinstance(dataset_array, device.BitPlanes) else dataset_array.shape[1]
    candidates = list(set(range(n_total)) - existing)
    if num_simulations > len(candidates):
        logging.warning(f"Requested {num_simulations} simulations, but only {len(candidates)} "
                        "unsimulated cells available.")
        num_simulations = len(candidates)
    if len(cells_to_simulate) == 0:
        cells_to_simulate = np.random.choice(candidates, num_simulations, replace=False)
    todo = [int(c) for c in cells_to_simulate
            if not os.path.exists(f'{trajectory_dir}/cell_{int(c)}_trajectory.csv')]
    if todo:
        traj = simulate_trajectories(network.nodes, dataset_array, todo)
        for cell, states in zip(todo, traj):
            with open(f'{trajectory_dir}/cell_{cell}_trajectory.csv', 'w') as fh:
                for node, row in zip(network.nodes, states):
                    fh.write(f'{node.name},{",".join(str(int(v)) for v in row)}\n')
    return cells_to_simulate


EUCLIDEAN_THRESHOLD = 5   # attractor_analysis.py:41


def _patterns_of(cell_trajectory_dict: dict, genes: list, downsample_factor: int):
    """uint16 [cells, genes] patterns (bit k = point k * factor) and the number of points."""
    names = list(cell_trajectory_dict.keys())
    first = np.asarray(cell_trajectory_dict[names[0]][genes[0]]).reshape(-1)
    points = len(first[::downsample_factor])
    if points > 12:
        raise ValueError("trajectory distance: more than 12 points per down-sampled series")
    traj = np.zeros((len(names), len(genes), len(first)), dtype=np.uint8)
    for m, cell in enumerate(names):
        for g, gene in enumerate(genes):
            series = np.asarray(cell_trajectory_dict[cell][gene]).reshape(-1)
            if series.size != first.size or not np.isin(series, (0, 1)).all():
                raise ValueError("trajectory distance: series must be 0/1 and of equal length")
            traj[m, g] = series
    torch = device._torch()
    pats = device.trajectory_patterns(torch.from_numpy(traj).cuda(), downsample_factor)
    return names, pats, points


def compute_dtw_distances(cell_trajectory_dict: dict, radius: int = 1, downsample_factor: int = 2) -> dict:
    """``{(cell1, cell2): total distance}`` for every pair in ``itertools.combinations`` order
    (reference: ``compute_dtw_distances``, attractor_analysis.py:347-377, a process pool over the
    pairs) from ONE ``scb_pattern_pair_distances`` call.  Every cell must hold the same genes (what
    ``simulate_cells`` writes); a pair without a shared gene is ``inf`` in the reference.
    The FastDTW branch (squared difference >= 5) follows a restatement of ``fastdtw`` 0.3.4: parity
    unpinned, see DESIGN.md."""
    names = list(cell_trajectory_dict.keys())
    if len(names) < 2:
        return {}
    genes = list(cell_trajectory_dict[names[0]].keys())
    if any(list(cell_trajectory_dict[c].keys()) != genes for c in names):
        raise ValueError("trajectory distance: every cell must hold the same genes in the same order")
    if not genes:
        return {(a, b): float("inf") for i, a in enumerate(names) for b in names[i + 1:]}
    names, pats, points = _patterns_of(cell_trajectory_dict, genes, downsample_factor)
    dist = device.pattern_pair_distances(pats, points, radius, EUCLIDEAN_THRESHOLD).cpu().numpy()
    return {(a, names[j]): int(dist[i, j]) for i, a in enumerate(names) for j in range(i + 1, len(names))}


def compute_dtw_distance_pair(cell1: str, cell2: str, cell_trajectory_dict: dict, radius: int = 1,
                              downsample_factor: int = 2):
    """``(cell1, cell2, total distance)`` (reference signature, attractor_analysis.py:304-345)."""
    pair = {cell1: cell_trajectory_dict[cell1], cell2: cell_trajectory_dict[cell2]}
    shared = [g for g in pair[cell1] if g in pair[cell2]]
    if not shared:
        return (cell1, cell2, float("inf"))
    pair = {c: {g: pair[c][g] for g in shared} for c in (cell1, cell2)}
    return (cell1, cell2, compute_dtw_distances(pair, radius, downsample_factor)[(cell1, cell2)])


def create_distance_matrix(dtw_distances: dict, file_names: list) -> np.ndarray:
    """Symmetric float matrix of the pair distances (attractor_analysis.py:381-405)."""
    index = {name: i for i, name in enumerate(file_names)}
    matrix = np.zeros((len(file_names), len(file_names)))
    for (a, b), total in dtw_distances.items():
        matrix[index[a], index[b]] = total
        matrix[index[b], index[a]] = total
    return matrix
