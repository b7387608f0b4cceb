"""Per-cell trajectory simulation, GPU-backed (reference: ``code/attractor_analysis.py:43-244``).

Only the simulation half of the reference module is on the hot path; DTW clustering and the
reporting code that follow it are out of scope.  ``vectorized_run_simulation`` keeps the
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
