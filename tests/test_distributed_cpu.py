"""CPU suite: cell sharding + integer all-reduce with world_size 2 over gloo.

The kernels cannot run here, so each rank computes its shard's joint counts with the NumPy
checker; what is under test is the product's sharding arithmetic (``distributed.word_range`` /
``shard_dense`` / ``shard_packed``) and ``all_reduce_sum`` — the pieces the GPU path uses
unchanged with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scbonita2_b200 import distributed as scd
from tests.helpers import np_counts
from tools import synth


def test_word_ranges_partition_and_align():
    for n_cells in (1, 31, 32, 33, 127, 128, 129, 1000, 3621, 100_000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            n_words = (n_cells + 31) // 32
            edges = [scd.word_range(n_cells, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n_words
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 <= a1
            for w0, w1 in edges:
                assert w0 % 4 == 0 or w0 == w1           # 16-byte aligned shard rows (or empty)
            cells = [scd.cell_range(n_cells, r, world) for r in range(world)]
            assert sum(c1 - c0 for c0, c1 in cells) == n_cells
            sizes = [w1 - w0 for w0, w1 in edges[:-1]]   # the last shard absorbs the ragged tail
            if sizes:
                assert max(sizes) - min(sizes) <= 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_nodes, n_cells, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        specs, planes = synth.synthetic_dataset(n_nodes, n_cells, seed, flip=0.1, relax_steps=2)
        dense = synth.planes_to_dense(planes, n_cells)
        # both shard views must describe the same cells
        local_dense = scd.shard_dense(dense, rank, world)
        local_words, local_cells = scd.shard_packed(planes, n_cells, rank, world)
        assert local_cells == local_dense.shape[1]
        assert (synth.planes_to_dense(local_words, local_cells) == local_dense).all()
        counts = np.stack([np_counts(local_dense, s.index, synth.spec_parent_rows(s)) for s in specs])
        t = torch.from_numpy(counts.copy())
        extra = torch.tensor([local_cells], dtype=torch.int64)
        scd.all_reduce_sum((t, extra))
        assert scd.world() == (rank, world)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), t.numpy())
        assert int(extra.item()) == n_cells
        with pytest.raises(TypeError):
            scd.all_reduce_sum(torch.zeros(2, dtype=torch.float32))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_nodes,n_cells,seed", [(9, 1000, 1), (5, 130, 2), (4, 33, 3)])
def test_sharded_counts_allreduce_gloo(tmp_path, n_nodes, n_cells, seed):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_nodes, n_cells, seed, str(tmp_path)), nprocs=world, join=True)
    specs, planes = synth.synthetic_dataset(n_nodes, n_cells, seed, flip=0.1, relax_steps=2)
    dense = synth.planes_to_dense(planes, n_cells)
    want = np.stack([np_counts(dense, s.index, synth.spec_parent_rows(s)) for s in specs])
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npy"))
        assert (got == want).all()       # every rank holds the full-dataset counts, bit-exact
