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


class PeerExchange:
    """The joint counts of the cell shards summed over peer memory instead of an all-reduce
    (``scb_peer_push_counts`` + ``scb_population_fitness_peers``, include/scbonita_b200.h 2c).

    Every rank owns a small region that all its peers map (CUDA IPC, NVLink); ``push(counts)``
    stores the rank's ``[G, 16]`` counts into its slot of every region, ``population_fitness``
    waits for the flags of all ranks and sums the slots while it builds its tables.  No host
    synchronisation, no library collective between the kernels; integer sums, so the result is
    bit-identical to the single-GPU one.  The same calls must run on every rank.
    """

    def __init__(self, n_groups: int, rank: int, world_size: int, region: int, peer_ptrs, owned_peers=()):
        import torch

        self.n_groups, self.rank, self.world_size = int(n_groups), int(rank), int(world_size)
        self.region = int(region)
        self._owned_peers = list(owned_peers)      # IPC mappings this object must close
        self.peer_table = torch.tensor([int(p) for p in peer_ptrs], dtype=torch.int64, device="cuda")
        self._closed = False

    # ---- constructors ---------------------------------------------------------------
    @staticmethod
    def _alloc(n_groups: int, world_size: int):
        import ctypes

        from . import _lib
        L = _lib.lib()
        nbytes = int(L.scb_peer_region_bytes(world_size, n_groups * 16))
        if nbytes <= 0:
            raise ValueError(f"peer exchange supports 1..8 ranks, got {world_size}")
        region = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(L.scb_peer_region_alloc(nbytes, ctypes.byref(region), handle), "scb_peer_region_alloc")
        return int(region.value), handle.raw

    @classmethod
    def create(cls, n_groups: int, group=None) -> "PeerExchange":
        """One exchange per rank of ``group`` (default process group): allocates the local region,
        trades IPC handles through the process group and maps every peer's region."""
        import ctypes

        import torch.distributed as dist

        from . import _lib
        rank, world_size = dist.get_rank(group), dist.get_world_size(group)
        region, handle = cls._alloc(n_groups, world_size)
        handles = [None] * world_size
        dist.all_gather_object(handles, handle, group=group)
        L = _lib.lib()
        ptrs, owned = [], []
        for r, h in enumerate(handles):
            if r == rank:
                ptrs.append(region)
                continue
            p = ctypes.c_void_p()
            _lib.check(L.scb_peer_region_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)),
                       "scb_peer_region_open")
            ptrs.append(int(p.value))
            owned.append(int(p.value))
        dist.barrier(group=group)
        return cls(n_groups, rank, world_size, region, ptrs, owned)

    @classmethod
    def local_group(cls, n_groups: int, world_size: int) -> list:
        """``world_size`` exchanges on ONE device wired to each other directly (no IPC): the same
        kernels and protocol, used by the single-GPU tests."""
        regions = [cls._alloc(n_groups, world_size)[0] for _ in range(world_size)]
        return [cls(n_groups, r, world_size, regions[r], regions) for r in range(world_size)]

    # ---- the two halves of the exchange ------------------------------------------------
    def push(self, counts):
        import torch

        from . import _lib
        from .device import _stream
        if counts.dtype != torch.int64 or counts.numel() != self.n_groups * 16 or not counts.is_contiguous():
            raise ValueError("counts must be a contiguous int64 [G, 16] tensor")
        _lib.check(_lib.lib().scb_peer_push_counts(counts.data_ptr(), self.n_groups * 16, self.region,
                                                   self.peer_table.data_ptr(), self.rank, self.world_size,
                                                   _stream()), "scb_peer_push_counts")

    def population_fitness(self, lut, active=None, per_node: bool = False, out=None):
        """As ``device.population_fitness`` on the counts summed over all ranks."""
        import torch

        from . import _lib
        from .device import _dev, _stream
        dev = self.peer_table.device
        d_lut = lut if (isinstance(lut, torch.Tensor) and lut.dtype == torch.uint8 and lut.is_cuda
                        and lut.is_contiguous()) else _dev(lut, torch.uint8, dev)
        n_ind, n_groups = int(d_lut.shape[0]), int(d_lut.shape[1])
        if n_groups != self.n_groups:
            raise ValueError("lut must be [P, G] with G == the exchange's group count")
        d_act = _dev(active, torch.uint8, dev) if active is not None else None
        if out is None:
            out = torch.empty(n_ind, dtype=torch.int64, device=dev)
        out_pg = torch.empty((n_ind, n_groups), dtype=torch.int64, device=dev) if per_node else None
        _lib.check(_lib.lib().scb_population_fitness_peers(
            self.region, self.rank, self.world_size, n_groups, n_ind, d_lut.data_ptr(),
            d_act.data_ptr() if d_act is not None else None, out.data_ptr(),
            out_pg.data_ptr() if out_pg is not None else None, _stream()), "scb_population_fitness_peers")
        return out, out_pg

    def check(self):
        """Raises if a kernel gave up waiting for a peer (synchronises the device)."""
        import ctypes

        from . import _lib
        err = ctypes.c_int(0)
        _lib.check(_lib.lib().scb_peer_error(self.region, ctypes.byref(err)), "scb_peer_error")
        if err.value:
            raise RuntimeError("peer exchange: a rank did not deliver its counts in time; the results of the "
                               "affected launches were poisoned (INT64_MIN)")

    def set_timeout(self, milliseconds: int):
        """How long a kernel waits for a peer's word before it gives up (0 = the default of 30 s)."""
        from . import _lib
        _lib.check(_lib.lib().scb_peer_set_timeout(self.region, int(milliseconds)), "scb_peer_set_timeout")

    def close(self):
        from . import _lib
        if self._closed:
            return
        self._closed = True
        L = _lib.lib()
        for p in self._owned_peers:
            L.scb_peer_region_close(p)
        L.scb_peer_region_free(self.region)
