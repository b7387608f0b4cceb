"""Cell sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Cells are independent for every kernel on the path; only small INTEGER vectors cross ranks
(joint counts ``[G,16]``, importance totals ``[N]``, missing-attractor counters).  Each rank owns
a contiguous, 128-bit aligned slice of the 32-cell words of every bitplane row, runs the same
kernels on its slice, and the integer outputs are summed with one all-reduce (NCCL over NVLink
on GPUs; gloo on CPU for the host-logic tests).  Integer sums are order-independent, so the
result is bit-identical at any GPU count.  The reference has no counterpart: it is single-host
``multiprocessing`` only (``importance_scores.py:238-253``).
"""
from __future__ import annotations

import numpy as np


def world():
    """(rank, world_size) of the default process group, (0, 1) if none is initialised."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def word_range(n_cells: int, rank: int, world_size: int) -> tuple:
    """Half-open range of 32-cell words owned by ``rank``: contiguous, multiples of 4 words
    (so every shard row starts 16-byte aligned), sizes differing by at most 4 words."""
    n_words = (n_cells + 31) // 32
    quads = (n_words + 3) // 4
    base, extra = divmod(quads, world_size)
    q0 = rank * base + min(rank, extra)
    q1 = q0 + base + (1 if rank < extra else 0)
    return min(4 * q0, n_words), min(4 * q1, n_words)


def cell_range(n_cells: int, rank: int, world_size: int) -> tuple:
    w0, w1 = word_range(n_cells, rank, world_size)
    return min(32 * w0, n_cells), min(32 * w1, n_cells)


def chunk_range(n_chunks: int, rank: int, world_size: int) -> tuple:
    """Chunk columns owned by ``rank`` on the chunked (majority-vote) path: every chunk is
    self-contained after the global permutation, so chunks shard like cells."""
    return cell_range(n_chunks, rank, world_size)


def shard_dense(matrix, rank: int, world_size: int):
    """Columns of a (genes x cells) matrix owned by ``rank`` (NumPy or scipy sparse)."""
    c0, c1 = cell_range(matrix.shape[1], rank, world_size)
    return matrix[:, c0:c1]


def shard_packed(packed: np.ndarray, n_cells: int, rank: int, world_size: int):
    """(words, n_cells_local) of host uint32 bitplanes owned by ``rank``."""
    w0, w1 = word_range(n_cells, rank, world_size)
    c0, c1 = cell_range(n_cells, rank, world_size)
    return np.ascontiguousarray(packed[:, w0:w1]), c1 - c0


def all_reduce_sum(tensors):
    """In-place SUM all-reduce of one tensor or a tuple of integer tensors; returns the input.
    No-op without an initialised process group.  On CUDA tensors this is ncclAllReduce on the
    current stream (enqueued behind the kernels that produced them, no host sync)."""
    import torch
    import torch.distributed as dist

    single = isinstance(tensors, torch.Tensor)
    items = (tensors,) if single else tuple(tensors)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in items:
            if t.dtype not in (torch.int64, torch.int32):
                raise TypeError("only integer tensors are reduced on this path (exactness)")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors if single else items


def sharded_rule_errors(nodes, matrix, rank=None, world_size=None):
    """``RuleDetermination.score_candidates`` on this rank's cell slice + all-reduce: every rank
    returns the full-dataset error lists.  ``matrix`` is the full (genes x cells) 0/1 matrix."""
    from . import device
    from .rule_determination import RuleDetermination

    if rank is None:
        rank, world_size = world()
    local = shard_dense(matrix, rank, world_size)
    planes = device.BitPlanes.from_dense(local)
    total_cells = matrix.shape[1]
    errors = RuleDetermination.score_candidates(nodes, planes, counts_hook=all_reduce_sum,
                                                count_override=total_cells)
    return errors


def sharded_importance(calculator):
    """Run ``CalculateImportanceScore`` built on this rank's cell slice; scores use the global
    integer totals."""
    return calculator.calculate_importance_scores(allreduce=all_reduce_sum)
