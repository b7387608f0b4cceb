"""CPU restatement ("port") of scBONITA2's hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product package ``scbonita2_b200`` never does; it fails loudly when its CUDA
library is missing.

Every function follows one reference routine (cited per function, paths relative to
``/root/reference/``) with the SAME algorithmic shape — Python ``eval`` per cell for rule
scoring, step-by-step synchronous simulation, all-pairs first-repeat search — so that timing
it is a fair stand-in for the reference's CPU cost.  Parity is PINNED: ``oracle/make_golden.py``
runs the unmodified reference (imported through ``oracle/ref_loader.py``) on seeded inputs and
commits inputs + outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays those
through this module on every run, with or without the reference mounted.

Inputs are plain Python/NumPy: a node is anything with ``name``, ``index``, ``predecessors``
(ordered dict row->name) and, for simulation, ``calculation_function`` (logic text).
"""
from __future__ import annotations

import random

import numpy as np

SLOT_NAMES = ("A", "B", "C", "D")


# --------------------------------------------------------------------------------------
# A4  rule scoring                                      code/rule_determination.py:509-582
# --------------------------------------------------------------------------------------
def calculate_error(node, logic: str, dataset: np.ndarray):
    """(difference, count) of ``logic`` for ``node`` over the columns of ``dataset``.

    ``dataset`` is the dense (rows x columns) 0/1 float matrix; slot i reads the row given by
    the i-th key of ``node.predecessors`` (:536-552); the rule text is evaluated by ``eval``
    once PER COLUMN (:557-562) and compared with row ``node.index`` (:531,577).
    """
    target = dataset[node.index]
    rows = [dataset[r] for r in list(node.predecessors)[:4]]
    names = SLOT_NAMES[:len(rows)]
    mismatches = 0
    for col in range(dataset.shape[1]):
        env = {nm: row[col] for nm, row in zip(names, rows)}
        value = eval(logic, {}, env)
        if value != target[col]:
            mismatches += 1
    return mismatches, int(dataset.shape[1])


# --------------------------------------------------------------------------------------
# A2  chunking                                          code/rule_determination.py:181-233
# --------------------------------------------------------------------------------------
def sparse_row_mean(count_ones: int, width: int) -> float:
    """Mean of a 0/1 row slice as ``scipy.sparse`` computes it: every stored 1 is first scaled
    by the float ``1/width`` and the scaled values are then accumulated one by one.  This is
    NOT ``count/width``: for width 12 and 6 ones it yields 0.49999999999999994."""
    step = 1.0 / width
    acc = 0.0
    for _ in range(count_ones):
        acc += step
    return acc


def chunk_threshold(width: int) -> int:
    """Smallest number of ones in a chunk of ``width`` cells whose float mean is >= 0.5."""
    for k in range(width + 1):
        if sparse_row_mean(k, width) >= 0.5:
            return k
    return width + 1


def chunk_plan(num_columns: int, num_chunks) -> tuple:
    """(permutation, n_chunks, chunk_size) exactly as ``chunk_data`` derives them (:191-209)."""
    num_chunks = int(num_chunks)
    np.random.seed(42)
    perm = np.random.permutation(num_columns)
    if num_chunks > num_columns:
        num_chunks = num_columns
    chunk_size = max(1, num_columns // num_chunks)
    return perm, num_chunks, chunk_size


def default_num_chunks(num_columns: int) -> int:
    """``round(C / 10)`` when C > 2000 else C (``rule_determination.py:26-33``)."""
    return round(num_columns / 10) if num_columns > 2000 else num_columns


def chunk_data(matrix01: np.ndarray, num_chunks) -> np.ndarray:
    """Majority-vote chunking of a dense 0/1 matrix -> float64 (rows, n_chunks)."""
    perm, n_chunks, size = chunk_plan(matrix01.shape[1], num_chunks)
    shuffled = np.asarray(matrix01)[:, perm]
    out = np.zeros((matrix01.shape[0], n_chunks))
    for k in range(n_chunks):
        block = shuffled[:, k * size:(k + 1) * size]
        width = block.shape[1]
        for r in range(block.shape[0]):
            ones = int(np.count_nonzero(block[r]))
            out[r, k] = 1.0 if sparse_row_mean(ones, width) >= 0.5 else 0.0
    return out


# --------------------------------------------------------------------------------------
# A9  synchronous simulation with knock-out / knock-in   code/importance_scores.py:47-111
# --------------------------------------------------------------------------------------
def _eval_bool(logic: str, env: dict):
    text = logic.replace("and", "&").replace("or", "|").replace("not", "~")
    return eval(text, {"__builtins__": {}}, {k: np.asarray(v).astype(bool) for k, v in env.items()})


def run_simulation(nodes, dataset_bool: np.ndarray, steps: int,
                   knockout_node=None, knockin_node=None) -> np.ndarray:
    """History (steps, nodes, cells) of bools.  History index 0 is the state after ONE update
    of the dataset columns; a knocked node is forced from index 0 on (:60-105)."""
    n_cells = dataset_bool.shape[1]
    history = np.zeros((steps, len(nodes), n_cells), dtype=bool)
    for step in range(steps):
        source = dataset_bool if step == 0 else history[step - 1]
        for pos, node in enumerate(nodes):
            if node.name == knockout_node:
                history[step, pos] = False
            elif node.name == knockin_node:
                history[step, pos] = True
            else:
                env = {nm: source[r] for nm, r in zip(SLOT_NAMES, list(node.predecessors)[:4])}
                history[step, pos] = _eval_bool(node.calculation_function, env)
    return history


# --------------------------------------------------------------------------------------
# A10 first-repeat attractor search                     code/importance_scores.py:114-143
# --------------------------------------------------------------------------------------
def find_attractor(states: np.ndarray):
    """``states``: (nodes, steps) of one cell.  Smallest i with an equal earlier column j
    (smallest such j) -> (j, i); (None, None) if no column repeats."""
    n_steps = states.shape[1]
    for i in range(n_steps):
        for j in range(i):
            if np.array_equal(states[:, i], states[:, j]):
                return j, i
    return None, None


def attractors_of(history: np.ndarray) -> dict:
    """{cell position: bool (nodes, lambda+1)} — the repeated state is kept at BOTH ends."""
    per_cell = np.transpose(history, (2, 1, 0))
    found = {}
    for cell in range(per_cell.shape[0]):
        j, i = find_attractor(per_cell[cell])
        if j is not None:
            found[cell] = per_cell[cell, :, j:i + 1]
    return found


# --------------------------------------------------------------------------------------
# A12 knock-out / knock-in scoring                      code/importance_scores.py:256-340
# --------------------------------------------------------------------------------------
def _aligned_xor_count(perturbed: np.ndarray, normal: np.ndarray) -> int:
    # align_attractors trims BOTH axes to the shorter extent (:325-340)
    rows = min(perturbed.shape[0], normal.shape[0])
    cols = min(perturbed.shape[1], normal.shape[1])
    return int(np.sum(perturbed[:rows, :cols] ^ normal[:rows, :cols]))


def importance_raw_counts(nodes, dataset_bool: np.ndarray, steps: int = 50):
    """Integer XOR totals per node (list, node order) + per-condition count of cells without a
    repeat (order: normal, KO_0, KI_0, KO_1, KI_1, ...).

    A cell contributes to a node only if BOTH its knock-out and knock-in runs repeat within
    ``steps`` (:278); if such a cell has no NORMAL attractor the reference raises ``KeyError``
    (:281) and so does this port.
    """
    normal = attractors_of(run_simulation(nodes, dataset_bool, steps))
    n_cells = dataset_bool.shape[1]
    missing = [n_cells - len(normal)]
    raw = []
    for node in nodes:
        ko = attractors_of(run_simulation(nodes, dataset_bool, steps, knockout_node=node.name))
        ki = attractors_of(run_simulation(nodes, dataset_bool, steps, knockin_node=node.name))
        missing += [n_cells - len(ko), n_cells - len(ki)]
        total = 0
        for cell in range(n_cells):
            if cell in ki and cell in ko:
                base = normal[cell]
                total += _aligned_xor_count(ki[cell], base)
                total += _aligned_xor_count(ko[cell], base)
        raw.append(total)
    return raw, missing


def sample_cells(n_columns: int, n_cells: int = 500) -> list:
    """The (Python ``random``) column sample of ``CalculateImportanceScore.__init__`` (:30-34);
    callers seed ``random`` first — the reference itself never does."""
    return random.sample(range(n_columns), k=min(n_cells, n_columns))


def importance_scores(nodes, dataset_bool: np.ndarray, steps: int = 50) -> dict:
    """{name: raw / max(raw)} (:313-317); ``ZeroDivisionError`` if every raw count is 0."""
    raw, _ = importance_raw_counts(nodes, dataset_bool, steps)
    top = max(raw)
    return {node.name: value / top for node, value in zip(nodes, raw)}


# --------------------------------------------------------------------------------------
# A13 per-cell trajectories                             code/attractor_analysis.py:43-99
# --------------------------------------------------------------------------------------
def cell_trajectory(nodes, cell_column, steps: int = 20) -> list:
    """``steps`` synchronous updates from one cell's column -> list[steps][nodes] of ints."""
    start = np.asarray(cell_column).astype(bool).reshape(-1, 1)
    history = run_simulation(nodes, start, steps)
    return history[:, :, 0].astype(int).tolist()


# --------------------------------------------------------------------------------------
# A7  historic GA fitness (bytecode only: code/__pycache__/updated_deap_class.cpython-36.pyc)
# --------------------------------------------------------------------------------------
def population_fitness(nodes, candidate_logics, selections, dataset: np.ndarray,
                       current_gen: int = 0, max_gen: int = 1):
    """Fitness of a population per the disassembled ``CustomDeap.fitness_calculation``.

    ``candidate_logics[n]`` is node n's ordered candidate list, ``selections[p][n]`` the list
    of candidate indices individual p switches on for node n.
    raw = sum(differences) / sum(counts); fitness = raw + 0.002 * (#selected rules) * gen/max_gen.
    PARITY UNPINNED by the reference for this function (no runnable source): it
    is pinned to A4 through the identity "sum of ``calculate_error`` over the selected rules".
    """
    raws, fits = [], []
    for individual in selections:
        diff_sum = 0
        count_sum = 0
        n_rules = 0
        for node, logics, chosen in zip(nodes, candidate_logics, individual):
            for ridx in chosen:
                d, c = calculate_error(node, logics[ridx], dataset)
                diff_sum += d
                count_sum += c
                n_rules += 1
        raw = diff_sum / count_sum
        raws.append(raw)
        fits.append((raw + 0.002 * n_rules * current_gen / max_gen,))
    return raws, fits


# --------------------------------------------------------------------------------------
# next-3  trajectory distance                          code/attractor_analysis.py:289-377
# --------------------------------------------------------------------------------------
EUCLIDEAN_THRESHOLD = 5   # attractor_analysis.py:41


def _fastdtw_reduce_by_half(x):
    return [(x[i] + x[i + 1]) / 2 for i in range(0, len(x) - len(x) % 2, 2)]


def _fastdtw_expand_window(path, len_x, len_y, radius):
    path_ = set(path)
    for i, j in path:
        for a in range(-radius, radius + 1):
            for b in range(-radius, radius + 1):
                path_.add((i + a, j + b))
    window_ = set()
    for i, j in path_:
        for a, b in ((i * 2, j * 2), (i * 2, j * 2 + 1), (i * 2 + 1, j * 2), (i * 2 + 1, j * 2 + 1)):
            window_.add((a, b))
    window = []
    start_j = 0
    for i in range(0, len_x):
        new_start_j = None
        for j in range(start_j, len_y):
            if (i, j) in window_:
                window.append((i, j))
                if new_start_j is None:
                    new_start_j = j
            elif new_start_j is not None:
                break
        start_j = new_start_j
    return window


def _fastdtw_dtw(x, y, window, dist):
    len_x, len_y = len(x), len(y)
    if window is None:
        window = [(i, j) for i in range(len_x) for j in range(len_y)]
    inf = float("inf")
    D = {(0, 0): (0, 0, 0)}
    for i, j in ((i + 1, j + 1) for i, j in window):
        dt = dist(x[i - 1], y[j - 1])
        # first minimum in the order (up, left, diagonal), like Python's min() over that tuple
        best = None
        for pi, pj in ((i - 1, j), (i, j - 1), (i - 1, j - 1)):
            cost = D.get((pi, pj), (inf,))[0] + dt
            if best is None or cost < best[0]:
                best = (cost, pi, pj)
        D[i, j] = best
    path = []
    i, j = len_x, len_y
    while not (i == j == 0):
        path.append((i - 1, j - 1))
        i, j = D[i, j][1], D[i, j][2]
    path.reverse()
    return D[len_x, len_y][0], path


def fastdtw(x, y, radius: int = 1, dist=lambda a, b: abs(a - b)):
    """Restatement of ``fastdtw.fastdtw`` (package ``fastdtw`` 0.3.4, ``environment.yml:103``; pure-Python
    module ``fastdtw/fastdtw.py``: Salvador & Chan's FastDTW -- halve both series by pair averages
    down to ``radius + 2`` points, exact DTW there, then project the path, widen it by ``radius``
    and run the windowed DTW one level up).  The package is NOT in the reference tree nor in this
    image: PARITY UNPINNED for this function (the recursion, the window construction and the
    up / left / diagonal tie order are restated from the published source; no golden vector).
    """
    x = [float(v) for v in np.asarray(x, dtype=float).reshape(-1)]
    y = [float(v) for v in np.asarray(y, dtype=float).reshape(-1)]

    def rec(xs, ys):
        min_time_size = radius + 2
        if len(xs) < min_time_size or len(ys) < min_time_size:
            return _fastdtw_dtw(xs, ys, None, dist)
        _, path = rec(_fastdtw_reduce_by_half(xs), _fastdtw_reduce_by_half(ys))
        window = _fastdtw_expand_window(path, len(xs), len(ys), radius)
        return _fastdtw_dtw(xs, ys, window, dist)

    return rec(x, y)


def dtw_distance_pair(traj1: dict, traj2: dict, radius: int = 1, downsample_factor: int = 2):
    """``compute_dtw_distance_pair`` (attractor_analysis.py:304-345) for two cells' ``{gene: series}``
    dicts: per shared gene, every ``downsample_factor``-th point; squared difference if it is below
    ``EUCLIDEAN_THRESHOLD``, else FastDTW with squared pointwise distance; the sum over the genes
    (``inf`` without a shared gene).  The squared-difference branch is plain NumPy (pinned by
    construction); the FastDTW branch inherits ``fastdtw``'s "parity unpinned".
    """
    total, any_gene = 0, False
    for gene, ts1 in traj1.items():
        if gene not in traj2:
            continue
        any_gene = True
        d1 = np.array(np.array(ts1)[::downsample_factor]).flatten()
        d2 = np.array(np.array(traj2[gene])[::downsample_factor]).flatten()
        squared = np.sum((d1 - d2) ** 2)
        if squared < EUCLIDEAN_THRESHOLD:
            total += squared
        else:
            total += fastdtw(d1, d2, radius, dist=lambda a, b: (a - b) ** 2)[0]
    return total if any_gene else float("inf")
