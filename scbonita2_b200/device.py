"""Device-side plumbing: cell bitplanes in HBM and thin typed wrappers over the C-ABI.

PyTorch is used for device memory, streams and (in ``distributed.py``) NCCL only; every
computation below is a call into ``libscbonita_b200.so``.  Tensors handed to the library are
``torch.int32`` (bitplane words — torch has no uint32 arithmetic and none is needed),
``torch.int64`` (counts) or ``torch.uint8`` (truth tables).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("scbonita2_b200 needs a CUDA device (there is no CPU fallback)")
    return torch


def _stream():
    return _torch().cuda.current_stream().cuda_stream


def padded_words(n_cells: int) -> int:
    """Words per bitplane row: ceil(n_cells/32) rounded up to a multiple of 4 (>= 4)."""
    return max(4, (((n_cells + 31) // 32) + 3) // 4 * 4)


def _dev(array, dtype, device=None):
    torch = _torch()
    if isinstance(array, torch.Tensor):
        return array.to(device=device or "cuda", dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(array), dtype=dtype).to(device or "cuda")


class BitPlanes:
    """uint32 cell bitplanes ``[n_rows][wpad]`` resident on one GPU.

    Row r is the gene with ``Node.index == r``; bit b of word w is cell ``32*w + b``; bits at
    and beyond ``n_cells`` are zero.
    """

    def __init__(self, words, n_rows: int, n_cells: int):
        self.words = words            # torch.int32 [n_rows, wpad], cuda
        self.n_rows = int(n_rows)
        self.n_cells = int(n_cells)
        self.wpad = int(words.shape[1])

    # ---- constructors ---------------------------------------------------------------
    @classmethod
    def from_dense(cls, matrix, device=None) -> "BitPlanes":
        """From a (rows x cells) 0/1 matrix: NumPy array or scipy sparse (any dtype)."""
        torch = _torch()
        if hasattr(matrix, "toarray"):
            matrix = matrix.toarray()
        dense = np.ascontiguousarray(np.asarray(matrix) != 0, dtype=np.uint8)
        if dense.ndim != 2:
            raise ValueError("expected a 2-D (genes x cells) matrix")
        n_rows, n_cells = dense.shape
        wpad = padded_words(n_cells)
        d_dense = torch.from_numpy(dense).to(device or "cuda")
        words = torch.empty((n_rows, wpad), dtype=torch.int32, device=d_dense.device)
        with torch.cuda.device(d_dense.device):
            _lib.check(_lib.lib().scb_pack_bitplanes(
                d_dense.data_ptr(), n_rows, n_cells, n_cells, words.data_ptr(), wpad, _stream()),
                "scb_pack_bitplanes")
        return cls(words, n_rows, n_cells)

    @classmethod
    def from_float(cls, matrix, threshold: float, device=None) -> "BitPlanes":
        """From a dense (rows x cells) float32/float64 expression matrix (NumPy array or CUDA
        tensor): bit = value > threshold, the reference's csr + ``sklearn.binarize`` recipe
        (``rule_inference.py:61-73``) in one pass, without the uint8 / csr intermediates."""
        torch = _torch()
        if not isinstance(matrix, torch.Tensor):
            matrix = np.ascontiguousarray(matrix)
            if matrix.dtype not in (np.float32, np.float64):
                matrix = matrix.astype(np.float64)
            matrix = torch.from_numpy(matrix)
        if matrix.ndim != 2 or matrix.dtype not in (torch.float32, torch.float64):
            raise ValueError("expected a 2-D float32/float64 (genes x cells) matrix")
        if threshold < 0:
            raise ValueError("Cannot binarize a sparse matrix with threshold < 0")   # as the reference
        d = matrix.to(device or "cuda").contiguous()
        n_rows, n_cells = int(d.shape[0]), int(d.shape[1])
        wpad = padded_words(n_cells)
        words = torch.empty((n_rows, wpad), dtype=torch.int32, device=d.device)
        fn = _lib.lib().scb_binarize_pack_f32 if d.dtype == torch.float32 else _lib.lib().scb_binarize_pack_f64
        with torch.cuda.device(d.device):
            _lib.check(fn(d.data_ptr(), n_rows, n_cells, n_cells, float(threshold), words.data_ptr(), wpad,
                          _stream()), "scb_binarize_pack")
        return cls(words, n_rows, n_cells)

    @classmethod
    def from_packed(cls, packed: np.ndarray, n_cells: int, device=None) -> "BitPlanes":
        """From host uint32 words ``[n_rows][>= ceil(n_cells/32)]`` (tail bits must be zero)."""
        torch = _torch()
        packed = np.ascontiguousarray(packed, dtype=np.uint32)
        n_rows, have = packed.shape
        n_words = (n_cells + 31) // 32
        if have < n_words:
            raise ValueError("packed rows are shorter than ceil(n_cells/32) words")
        wpad = padded_words(n_cells)
        host = np.zeros((n_rows, wpad), dtype=np.uint32)
        host[:, :n_words] = packed[:, :n_words]
        words = torch.from_numpy(host.view(np.int32)).to(device or "cuda")
        return cls(words, n_rows, n_cells)

    # ---- views ----------------------------------------------------------------------
    def to_dense(self) -> np.ndarray:
        torch = _torch()
        out = torch.empty((self.n_rows, self.n_cells), dtype=torch.uint8, device=self.words.device)
        with torch.cuda.device(self.words.device):
            _lib.check(_lib.lib().scb_unpack_bitplanes(
                self.words.data_ptr(), self.n_rows, self.n_cells, self.wpad, out.data_ptr(),
                self.n_cells, _stream()), "scb_unpack_bitplanes")
        return out.cpu().numpy()

    def to_packed(self) -> np.ndarray:
        n_words = (self.n_cells + 31) // 32
        return self.words.cpu().numpy().view(np.uint32)[:, :n_words].copy()

    def word_slice(self, w0: int, w1: int) -> "BitPlanes":
        """Cells [32*w0, 32*w1) as an independent BitPlanes (used for cell sharding)."""
        torch = _torch()
        w1 = min(w1, (self.n_cells + 31) // 32)
        n_cells = max(0, min(self.n_cells, 32 * w1) - 32 * w0)
        wpad = padded_words(n_cells)
        words = torch.zeros((self.n_rows, wpad), dtype=torch.int32, device=self.words.device)
        if w1 > w0:
            words[:, :w1 - w0] = self.words[:, w0:w1]
        return BitPlanes(words, self.n_rows, n_cells)

    # ---- chunking -------------------------------------------------------------------
    def chunk_majority(self, perm: np.ndarray, n_chunks: int, chunk_size: int,
                       min_ones: int) -> "BitPlanes":
        torch = _torch()
        d_perm = _dev(np.asarray(perm[:n_chunks * chunk_size], dtype=np.int32), torch.int32,
                      self.words.device)
        wout = padded_words(n_chunks)
        out = torch.empty((self.n_rows, wout), dtype=torch.int32, device=self.words.device)
        with torch.cuda.device(self.words.device):
            _lib.check(_lib.lib().scb_chunk_majority(
                self.words.data_ptr(), self.n_rows, self.wpad, d_perm.data_ptr(), n_chunks,
                chunk_size, min_ones, out.data_ptr(), wout, _stream()), "scb_chunk_majority")
        return BitPlanes(out, self.n_rows, n_chunks)


# --------------------------------------------------------------------------------------
def _host_i32(values, shape):
    torch = _torch()
    if isinstance(values, torch.Tensor):
        values = values.cpu().numpy()
    return np.ascontiguousarray(np.asarray(values, dtype=np.int32).reshape(shape))


class RuleCountsPlan:
    """Device-resident metadata, scratch and output of ``scb_rule_counts`` for a fixed set of
    groups, so that repeated passes (GA generations, benchmark steps) enqueue two kernels and
    nothing else: no host->device copies, no allocation, no synchronisation."""

    def __init__(self, n_rows: int, target_row, parent_row, device=None):
        torch = _torch()
        target = _host_i32(target_row, (-1,))
        parents = _host_i32(parent_row, (-1, 3))
        if parents.shape[0] != target.size:
            raise ValueError("parent_row must be [G, 3]")
        if target.size and (target.min() < 0 or target.max() >= n_rows or parents.min() < -1
                            or parents.max() >= n_rows):
            raise IndexError("rule_counts: row index out of range")
        dev = torch.device(device or "cuda")
        self.n_rows = int(n_rows)
        self.n_groups = int(target.size)
        self.target = torch.from_numpy(target).to(dev)
        self.parents = torch.from_numpy(parents).to(dev)
        L = _lib.lib()
        with torch.cuda.device(dev):
            self.scratch_bytes = int(L.scb_rule_counts_scratch_bytes(self.n_rows, self.n_groups))
            self.plan_bytes = int(L.scb_rule_counts_plan_bytes(self.n_rows, self.n_groups))
            # zeroed ONCE: the kernel leaves its scratch zeroed, so run() enqueues no memset
            self.scratch = torch.zeros(max(self.scratch_bytes, 16), dtype=torch.uint8, device=dev)
            self.plan = torch.empty(max(self.plan_bytes, 16), dtype=torch.uint8, device=dev)
            # the cell-independent tables (arity order, row offsets, row owners) are built once
            _lib.check(L.scb_rule_counts_plan_build(self.n_rows, self.n_groups, self.target.data_ptr(),
                                                    self.parents.data_ptr(), self.plan.data_ptr(),
                                                    self.plan_bytes, _stream()), "scb_rule_counts_plan_build")
        self.counts = torch.empty((self.n_groups, 16), dtype=torch.int64, device=dev)
        self.n_untargeted = int(self.n_rows - len(set(int(t) for t in target))) if target.size else self.n_rows
        self.fused_supported = bool(self.n_groups) and bool(
            L.scb_rule_sums_supported(self.n_rows, self.n_groups, self.n_untargeted))

    def _flags(self, chained: bool) -> int:
        return _lib.SCB_COUNTS_SCRATCH_CLEAN | (_lib.SCB_COUNTS_CHAINED if chained else 0)

    def sums_run(self, planes: BitPlanes, exchange=None, chained: bool = False):
        """First half of the fused GA step: ``scb_rule_sums_run`` -- one pass over the cells that leaves
        the SUBSET SUMS of every group in the plan's scratch (and, with ``exchange``, pushes them to the
        peers).  Must be followed by exactly one ``fitness_from_sums`` on this plan."""
        torch = _torch()
        if not self.fused_supported:
            raise ValueError("more groups than the fused step holds (about 660: its fitness kernel keeps a copy "
                             "of the sums in shared memory): use run() + population_fitness()")
        if planes.n_rows != self.n_rows:
            raise ValueError("plan was built for a different number of rows")
        if exchange is not None and exchange.n_groups != self.n_groups:
            raise ValueError("exchange was built for a different number of groups")
        with torch.cuda.device(self.counts.device):
            _lib.check(_lib.lib().scb_rule_sums_run(
                planes.words.data_ptr(), planes.n_rows, planes.wpad, planes.n_cells, self.n_groups, self.n_untargeted,
                self.plan.data_ptr(), self.scratch.data_ptr(), self.scratch_bytes, self._flags(chained),
                exchange.peer_table.data_ptr() if exchange is not None else None,
                exchange.rank if exchange is not None else 0, exchange.world_size if exchange is not None else 1,
                _stream()), "scb_rule_sums_run")
        self._sums_cells = planes.n_cells

    def fitness_from_sums(self, lut, out=None, exchange=None, active=None, per_node: bool = False):
        """Second half: ``scb_population_fitness_sums`` -- inverts the sums while it builds its tables,
        evaluates the population and leaves the scratch cleared for the next pass."""
        torch = _torch()
        dev = self.counts.device
        if isinstance(lut, torch.Tensor) and lut.dtype == torch.uint8 and lut.device == dev and lut.is_contiguous():
            d_lut = lut
        else:
            d_lut = _dev(lut if isinstance(lut, torch.Tensor) else np.asarray(lut, dtype=np.uint8), torch.uint8, dev)
        n_ind = int(d_lut.shape[0])
        if int(d_lut.shape[1]) != self.n_groups:
            raise ValueError("lut must be [P, G] with G == the plan's group count")
        d_act = _dev(active, torch.uint8, dev) if active is not None else None
        if out is None:
            out = torch.empty(n_ind, dtype=torch.int64, device=dev)
        out_pg = torch.empty((n_ind, self.n_groups), dtype=torch.int64, device=dev) if per_node else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().scb_population_fitness_sums(
                self.plan.data_ptr(), self.scratch.data_ptr(), self.n_rows, self.n_groups, self._sums_cells,
                exchange.region if exchange is not None else None,
                exchange.rank if exchange is not None else 0, exchange.world_size if exchange is not None else 1,
                n_ind, d_lut.data_ptr(), d_act.data_ptr() if d_act is not None else None, out.data_ptr(),
                out_pg.data_ptr() if out_pg is not None else None, _stream()), "scb_population_fitness_sums")
        return out, out_pg

    def fitness_step(self, planes: BitPlanes, lut, out=None, exchange=None, active=None, per_node: bool = False,
                     chained: bool = False):
        """One GA-fitness step, fused: ``sums_run`` + ``fitness_from_sums`` -- two kernel launches and
        nothing else; no ticket, no last-CTA finalize, no memset between or after them.  With
        ``exchange`` the sums of the cell shards travel through the peers' exchange regions.
        ``chained=True`` is the caller's word that the kernel enqueued before this call wrote neither the
        bitplanes nor the plan (e.g. it is the previous ``fitness_step``): the pass then overlaps that
        kernel's tail.  Same numbers as ``run`` + ``population_fitness``."""
        self.sums_run(planes, exchange=exchange, chained=chained)
        return self.fitness_from_sums(lut, out=out, exchange=exchange, active=active, per_node=per_node)

    def run(self, planes: BitPlanes, exchange=None, chained: bool = False):
        """Enqueue the pass on the current stream; returns the (reused) int64 [G, 16] tensor.
        With ``exchange`` (``distributed.PeerExchange``) the kernel's last CTA also pushes the counts
        to every peer's exchange region (``scb_rule_counts_run_push``).  ``chained``: see ``fitness_step``."""
        torch = _torch()
        if planes.n_rows != self.n_rows:
            raise ValueError("plan was built for a different number of rows")
        if exchange is not None:
            if exchange.n_groups != self.n_groups:
                raise ValueError("exchange was built for a different number of groups")
            with torch.cuda.device(self.counts.device):
                _lib.check(_lib.lib().scb_rule_counts_run_push(
                    planes.words.data_ptr(), planes.n_rows, planes.wpad, planes.n_cells, self.n_groups,
                    self.plan.data_ptr(), self.scratch.data_ptr(), self.scratch_bytes,
                    self.counts.data_ptr(), self._flags(chained), exchange.peer_table.data_ptr(),
                    exchange.rank, exchange.world_size, _stream()), "scb_rule_counts_run_push")
            return self.counts
        with torch.cuda.device(self.counts.device):
            _lib.check(_lib.lib().scb_rule_counts_run(
                planes.words.data_ptr(), planes.n_rows, planes.wpad, planes.n_cells, self.n_groups,
                self.plan.data_ptr(), self.scratch.data_ptr(), self.scratch_bytes,
                self.counts.data_ptr(), self._flags(chained), _stream()), "scb_rule_counts_run")
        return self.counts


def rule_counts(planes: BitPlanes, target_row, parent_row):
    """int64 ``[G, 16]`` joint counts (device tensor) for G (target, <=3 parents) groups."""
    plan = RuleCountsPlan(planes.n_rows, target_row, parent_row, planes.words.device)
    return plan.run(planes)


def rule_error_table(counts, task_group, task_lut):
    """int64 ``[T]`` mismatch counts (device tensor) of T (group, truth table) tasks."""
    torch = _torch()
    dev = counts.device
    h_group = _host_i32(task_group, (-1,))
    n_tasks = int(h_group.size)
    if n_tasks and (h_group.min() < 0 or h_group.max() >= counts.shape[0]):
        raise IndexError("rule_error_table: group index out of range")
    group = torch.from_numpy(h_group).to(dev)
    lut = _dev(np.asarray(task_lut, dtype=np.uint8).reshape(-1), torch.uint8, dev)
    if lut.numel() != n_tasks:
        raise ValueError("task_group and task_lut differ in length")
    out = torch.empty(n_tasks, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().scb_rule_error_table(counts.data_ptr(), int(counts.shape[0]), n_tasks,
                                                   group.data_ptr(), lut.data_ptr(),
                                                   out.data_ptr(), _stream()),
                   "scb_rule_error_table")
    return out


def rule_error_table_direct(planes: BitPlanes, target_row, parent_row, lut):
    torch = _torch()
    dev = planes.words.device
    target = _dev(np.asarray(target_row, dtype=np.int32).reshape(-1), torch.int32, dev)
    parents = _dev(np.asarray(parent_row, dtype=np.int32).reshape(-1, 3), torch.int32, dev)
    d_lut = _dev(np.asarray(lut, dtype=np.uint8).reshape(-1), torch.uint8, dev)
    n_tasks = int(target.numel())
    out = torch.empty(n_tasks, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().scb_rule_error_table_direct(
            planes.words.data_ptr(), planes.n_rows, planes.wpad, planes.n_cells, n_tasks,
            target.data_ptr(), parents.data_ptr(), d_lut.data_ptr(), out.data_ptr(), _stream()),
            "scb_rule_error_table_direct")
    return out


def population_fitness(counts, lut, active=None, per_node: bool = False, out=None):
    """(diff[P], diff_pg[P, G] | None) device tensors for a population of truth-table rows.
    With ``lut`` already a device uint8 tensor and ``out`` a preallocated int64 [P] tensor the
    call enqueues one kernel and nothing else."""
    torch = _torch()
    dev = counts.device
    if isinstance(lut, torch.Tensor) and lut.dtype == torch.uint8 and lut.device == dev and lut.is_contiguous():
        d_lut = lut
    else:
        d_lut = _dev(lut if isinstance(lut, torch.Tensor) else np.asarray(lut, dtype=np.uint8), torch.uint8, dev)
    n_ind, n_groups = int(d_lut.shape[0]), int(d_lut.shape[1])
    if n_groups != counts.shape[0]:
        raise ValueError("lut must be [P, G] with G == counts.shape[0]")
    d_act = None
    if active is not None:
        d_act = _dev(active, torch.uint8, dev)
        if tuple(d_act.shape) != (n_ind, n_groups):
            raise ValueError("active must be [P, G]")
    if out is None:
        out = torch.empty(n_ind, dtype=torch.int64, device=dev)
    out_pg = torch.empty((n_ind, n_groups), dtype=torch.int64, device=dev) if per_node else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().scb_population_fitness(
            counts.data_ptr(), n_groups, n_ind, d_lut.data_ptr(),
            d_act.data_ptr() if d_act is not None else None, out.data_ptr(),
            out_pg.data_ptr() if out_pg is not None else None, _stream()),
            "scb_population_fitness")
    return out, out_pg


def bitstring_fitness(task_diff, bits):
    """(diff[P] int64, nsel[P] int32) device tensors for 0/1 strings over the task table."""
    torch = _torch()
    dev = task_diff.device
    d_bits = _dev(bits if isinstance(bits, torch.Tensor) else np.asarray(bits, dtype=np.uint8), torch.uint8, dev)
    n_ind, n_tasks = int(d_bits.shape[0]), int(d_bits.shape[1])
    if n_tasks != task_diff.numel():
        raise ValueError("bits must be [P, T] with T == len(task_diff)")
    out = torch.empty(n_ind, dtype=torch.int64, device=dev)
    nsel = torch.empty(n_ind, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().scb_bitstring_fitness(task_diff.data_ptr(), n_tasks, n_ind, d_bits.data_ptr(),
                                                    out.data_ptr(), nsel.data_ptr(), _stream()),
                   "scb_bitstring_fitness")
    return out, nsel


LAST_SWEEP_STATS = {}


def bind_host_to_device(index: int = None) -> bool:
    """Pin this process to the CPU cores that are local to GPU ``index`` (default: the current device), so that
    pinned host buffers allocated afterwards live on that NUMA node: a host-to-device copy from the other socket
    runs at about half the PCIe rate.  Uses NVML (``nvidia_ml_py``); returns False, changing nothing, when NVML or
    ``sched_setaffinity`` is unavailable or the query fails."""
    import os
    try:
        import pynvml
        torch = _torch()
        if index is None:
            index = torch.cuda.current_device()
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(index)
        handle = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                handle = None
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def ko_ki_sweep(planes: BitPlanes, net_parent, net_lut, steps: int = 50, out=None, stats: bool = False):
    """Run the 2N+1-condition sweep.  Returns device tensors (raw[N], missing[2N+1], keyerror[1]);
    pass ``out`` (the same triple) to accumulate several cell shards into one result."""
    torch = _torch()
    dev = planes.words.device
    n = planes.n_rows
    h_parents = _host_i32(net_parent, (-1, 3))
    d_lut = _dev(np.asarray(net_lut, dtype=np.uint8).reshape(-1), torch.uint8, dev)
    if h_parents.shape[0] != n or d_lut.numel() != n:
        raise ValueError("the sweep needs one (parents, lut) entry per bitplane row")
    if h_parents.min() < -1 or h_parents.max() >= n:
        raise IndexError("ko_ki_sweep: parent row out of range")
    parents = torch.from_numpy(h_parents).to(dev)
    if out is None:
        out = (torch.zeros(n, dtype=torch.int64, device=dev),
               torch.zeros(2 * n + 1, dtype=torch.int64, device=dev),
               torch.zeros(1, dtype=torch.int64, device=dev))
    L = _lib.lib()
    with torch.cuda.device(dev):
        nbytes = int(L.scb_ko_ki_sweep_scratch_bytes(n, planes.wpad, steps))
        scratch = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        _lib.check(L.scb_ko_ki_sweep(planes.words.data_ptr(), n, planes.wpad, planes.n_cells,
                                     parents.data_ptr(), d_lut.data_ptr(), steps,
                                     scratch.data_ptr(), nbytes, out[0].data_ptr(),
                                     out[1].data_ptr(), out[2].data_ptr(), _stream()),
                   "scb_ko_ki_sweep")
        if stats:  # diagnostic only (synchronises): deferred (word, node) pairs and ring-depth flag
            tail = scratch[nbytes - 16 - 8 * n:nbytes].view(torch.int32).cpu().numpy()
            LAST_SWEEP_STATS.update(rechecked_cells=int(tail[:n].sum()), rechecked_max_per_node=int(tail[:n].max()),
                                    deferred_cells=int(tail[n:2 * n].sum()),
                                    deferred_max_per_node=int(tail[n:2 * n].max()),
                                    deep_ring=bool(tail[2 * n + 1] & 1), fast_steps=int(tail[2 * n + 3]),
                                    cell_node_pairs=int(planes.n_cells) * n)
    return out


def trajectories(planes: BitPlanes, cell_idx, net_parent, net_lut, steps: int = 20,
                 forced_node: int = -1, forced_value: int = 0):
    """uint8 ``[M, N, steps]`` device tensor of states after 1..steps updates; optionally with one
    node knocked out (forced 0) or in (forced 1) from the first update on."""
    torch = _torch()
    dev = planes.words.device
    n = planes.n_rows
    idx = _dev(np.asarray(cell_idx, dtype=np.int64).reshape(-1), torch.int64, dev)
    if idx.numel() and (int(idx.min().item()) < 0 or int(idx.max().item()) >= planes.n_cells):
        raise IndexError("trajectories: cell index out of range")
    parents = _dev(np.asarray(net_parent, dtype=np.int32).reshape(-1, 3), torch.int32, dev)
    d_lut = _dev(np.asarray(net_lut, dtype=np.uint8).reshape(-1), torch.uint8, dev)
    out = torch.empty((int(idx.numel()), n, steps), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().scb_trajectories_forced(
            planes.words.data_ptr(), n, planes.wpad, planes.n_cells, idx.data_ptr(), int(idx.numel()),
            parents.data_ptr(), d_lut.data_ptr(), steps, int(forced_node), int(forced_value),
            out.data_ptr(), _stream()), "scb_trajectories_forced")
    return out


# --------------------------------------------------------------------------------------
_DTW_TABLES = {}


def dtw_table(bits: int, radius: int = 1, threshold: int = 5, device=None):
    """Device copy of the per-gene distance table over two ``bits``-point 0/1 patterns
    (``scb_dtw_table_build``; built once per (bits, radius, threshold, device))."""
    torch = _torch()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    key = (bits, radius, threshold, str(dev))
    if key not in _DTW_TABLES:
        host = np.zeros((1 << bits, 1 << bits), dtype=np.uint8)
        _lib.check(_lib.lib().scb_dtw_table_build(bits, radius, threshold, host.ctypes.data), "scb_dtw_table_build")
        _DTW_TABLES[key] = torch.from_numpy(host).to(dev)
    return _DTW_TABLES[key]


def trajectory_patterns(traj, factor: int = 2):
    """uint16 ``[M, N]`` patterns of a uint8 ``[M, N, steps]`` trajectory tensor: bit k = point
    ``k * factor`` (the reference's ``downsample``, attractor_analysis.py:289-291)."""
    torch = _torch()
    if traj.dtype != torch.uint8 or traj.dim() != 3:
        raise ValueError("trajectories must be a uint8 [cells, nodes, steps] tensor")
    traj = traj.contiguous()
    m, n, steps = (int(v) for v in traj.shape)
    out = torch.empty((m, n), dtype=torch.int16, device=traj.device)   # bit patterns; torch has no uint16 arithmetic
    with torch.cuda.device(traj.device):
        _lib.check(_lib.lib().scb_trajectory_patterns(traj.data_ptr(), m, n, steps, factor, out.data_ptr(), _stream()),
                   "scb_trajectory_patterns")
    return out


def pattern_pair_distances(patterns, bits: int, radius: int = 1, threshold: int = 5):
    """int32 ``[M, M]`` symmetric matrix: entry (a, b), a <= b, is the reference's
    ``compute_dtw_distance_pair(cell a, cell b)`` total over the genes."""
    torch = _torch()
    m, n = (int(v) for v in patterns.shape)
    table = dtw_table(bits, radius, threshold, patterns.device)
    out = torch.zeros((m, m), dtype=torch.int32, device=patterns.device)
    with torch.cuda.device(patterns.device):
        _lib.check(_lib.lib().scb_pattern_pair_distances(patterns.data_ptr(), m, n, table.data_ptr(), bits,
                                                         out.data_ptr(), _stream()), "scb_pattern_pair_distances")
    return out


class HostWorkspace:
    """The host-buffer ("end-to-end") entry points: NumPy in, NumPy out, no torch."""

    def __init__(self):
        import ctypes

        self._ct = ctypes
        self._L = _lib.lib()
        handle = ctypes.c_void_p()
        _lib.check(self._L.scb_workspace_create(ctypes.byref(handle)), "scb_workspace_create")
        self._h = handle
        self.n_rows = 0
        self.n_groups = 0

    def close(self):
        if self._h is not None:
            self._L.scb_workspace_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ptr(self, arr, ctype):
        return arr.ctypes.data_as(self._ct.POINTER(ctype))

    def load_dense(self, dense01: np.ndarray):
        ct = self._ct
        dense01 = np.ascontiguousarray(dense01, dtype=np.uint8)
        _lib.check(self._L.scb_host_load_dense(self._h, self._ptr(dense01, ct.c_uint8),
                                               dense01.shape[0], dense01.shape[1]),
                   "scb_host_load_dense")
        self.n_rows = dense01.shape[0]

    def load_planes(self, packed: np.ndarray, n_cells: int):
        ct = self._ct
        n_words = (n_cells + 31) // 32
        packed = np.ascontiguousarray(packed[:, :n_words], dtype=np.uint32)
        _lib.check(self._L.scb_host_load_planes(self._h, self._ptr(packed, ct.c_uint32),
                                                packed.shape[0], n_cells), "scb_host_load_planes")
        self.n_rows = packed.shape[0]

    def rule_counts(self, target_row, parent_row, want: bool = True):
        ct = self._ct
        target = np.ascontiguousarray(target_row, dtype=np.int32).reshape(-1)
        parents = np.ascontiguousarray(parent_row, dtype=np.int32).reshape(-1, 3)
        out = np.empty((target.size, 16), dtype=np.int64) if want else None
        _lib.check(self._L.scb_host_rule_counts(
            self._h, target.size, self._ptr(target, ct.c_int32), self._ptr(parents, ct.c_int32),
            self._ptr(out, ct.c_int64) if want else None), "scb_host_rule_counts")
        self.n_groups = target.size
        return out

    def rule_error_table(self, task_group, task_lut):
        ct = self._ct
        group = np.ascontiguousarray(task_group, dtype=np.int32).reshape(-1)
        lut = np.ascontiguousarray(task_lut, dtype=np.uint8).reshape(-1)
        out = np.empty(group.size, dtype=np.int64)
        _lib.check(self._L.scb_host_rule_error_table(
            self._h, group.size, self._ptr(group, ct.c_int32), self._ptr(lut, ct.c_uint8),
            self._ptr(out, ct.c_int64)), "scb_host_rule_error_table")
        return out

    def population_fitness(self, lut, active=None, per_node: bool = False):
        ct = self._ct
        lut = np.ascontiguousarray(lut, dtype=np.uint8)
        if lut.ndim != 2 or lut.shape[1] != self.n_groups:
            raise ValueError("lut must be [P, G]")
        act = np.ascontiguousarray(active, dtype=np.uint8) if active is not None else None
        out = np.empty(lut.shape[0], dtype=np.int64)
        out_pg = np.empty(lut.shape, dtype=np.int64) if per_node else None
        _lib.check(self._L.scb_host_population_fitness(
            self._h, lut.shape[0], self._ptr(lut, ct.c_uint8),
            self._ptr(act, ct.c_uint8) if act is not None else None,
            self._ptr(out, ct.c_int64), self._ptr(out_pg, ct.c_int64) if per_node else None),
            "scb_host_population_fitness")
        return out, out_pg

    def ko_ki_sweep(self, net_parent, net_lut, steps: int = 50):
        ct = self._ct
        parents = np.ascontiguousarray(net_parent, dtype=np.int32).reshape(-1, 3)
        lut = np.ascontiguousarray(net_lut, dtype=np.uint8).reshape(-1)
        n = self.n_rows
        if parents.shape[0] != n or lut.size != n:
            raise ValueError("the sweep needs one (parents, lut) entry per matrix row")
        raw = np.empty(n, dtype=np.int64)
        missing = np.empty(2 * n + 1, dtype=np.int64)
        keyerr = np.empty(1, dtype=np.int64)
        _lib.check(self._L.scb_host_ko_ki_sweep(
            self._h, self._ptr(parents, ct.c_int32), self._ptr(lut, ct.c_uint8), steps,
            self._ptr(raw, ct.c_int64), self._ptr(missing, ct.c_int64),
            self._ptr(keyerr, ct.c_int64)), "scb_host_ko_ki_sweep")
        return raw, missing, keyerr
