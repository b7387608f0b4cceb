"""GPU parity at BASELINE.json's full sizes (H1M: 200 nodes x 1,000,000 cells) through
size-independent properties, plus mid-size exact checks against the C oracle."""
import numpy as np
import pytest

from oracle import cport
from tests.helpers import nodes_of, np_counts, np_diff
from tools import synth

pytestmark = pytest.mark.gpu

N_NODES, N_CELLS = 200, 1_000_000


@pytest.fixture(scope="module")
def h1m(cuda):
    from scbonita2_b200 import device
    specs = synth.random_network(N_NODES, 0)
    start = synth.random_planes(N_NODES, N_CELLS, 1000)
    planes_np = synth.relax_and_flip(specs, start, N_CELLS, 0, relax_steps=8, flip=0.05)
    planes = device.BitPlanes.from_packed(planes_np, N_CELLS)
    targets = [s.index for s in specs]
    parents = [synth.spec_parent_rows(s) for s in specs]
    return specs, planes_np, planes, targets, parents


def test_h1m_pack_round_trip_and_counts(h1m):
    from scbonita2_b200 import device
    specs, planes_np, planes, targets, parents = h1m
    assert (planes.to_packed() == planes_np).all()
    counts = device.rule_counts(planes, targets, parents)
    c = counts.cpu().numpy()
    assert (c.sum(axis=1) == N_CELLS).all() and (c >= 0).all()
    # exact check of a few nodes against plain NumPy on the unpacked rows
    dense = synth.planes_to_dense(planes_np, N_CELLS)
    for g in (0, 57, 199):
        assert (c[g] == np_counts(dense, targets[g], parents[g])).all()
    # repacking the dense matrix on the GPU gives the same planes
    assert (device.BitPlanes.from_dense(dense).to_packed() == planes_np).all()
    # shard additivity at full size (3 uneven shards)
    n_words = planes_np.shape[1]
    parts = 0
    for lo, hi in [(0, 10_000), (10_000, 10_004), (10_004, n_words)]:
        parts = parts + device.rule_counts(planes.word_slice(lo, hi), targets, parents)
    assert (parts == counts).all()
    # cell order is irrelevant: reversing the words leaves every count unchanged
    rev = device.BitPlanes.from_packed(np.ascontiguousarray(planes_np[:, ::-1]), n_words * 32)
    c_rev = device.rule_counts(rev, targets, parents).cpu().numpy()
    c_rev[:, 0] -= n_words * 32 - N_CELLS          # the padding cells read (A,B,C,T) = 0
    assert (c_rev == c).all()


def test_h1m_population_fitness_properties(h1m):
    from scbonita2_b200 import device
    specs, planes_np, planes, targets, parents = h1m
    counts = device.rule_counts(planes, targets, parents)
    c = counts.cpu().numpy()
    rng = np.random.default_rng(3)
    lut = rng.integers(0, 256, (4096, N_NODES)).astype(np.uint8)
    diff, diff_pg = device.population_fitness(counts, lut, per_node=True)
    diff, diff_pg = diff.cpu().numpy(), diff_pg.cpu().numpy()
    assert (diff == diff_pg.sum(axis=1)).all()
    for p, n in [(0, 0), (17, 99), (4095, 199), (2048, 57)]:
        assert diff_pg[p, n] == np_diff(c[n], int(lut[p, n]))
    # complementing every truth table turns matches into mismatches
    diff_c, _ = device.population_fitness(counts, (~lut).astype(np.uint8))
    assert (diff + diff_c.cpu().numpy() == N_NODES * N_CELLS).all()
    # the brute-force kernel (LUT per word, no counts) agrees on a sample of (individual, node) tasks
    idx_p, idx_n = rng.integers(0, 4096, 64), rng.integers(0, N_NODES, 64)
    direct = device.rule_error_table_direct(planes, np.asarray(targets)[idx_n], np.asarray(parents)[idx_n],
                                            lut[idx_p, idx_n]).cpu().numpy()
    assert (direct == diff_pg[idx_p, idx_n]).all()


def test_sweep_midsize_exact_vs_c_oracle(cuda):
    from scbonita2_b200 import device
    from scbonita2_b200.importance_scores import compile_network
    for n_nodes, n_cells, seed, kw in [(60, 3000, 1, dict(p_inhibit=0.2)), (40, 2500, 2, dict(p_inhibit=0.5)),
                                       (200, 700, 3, dict(p_inhibit=0.2))]:
        specs, planes_np = synth.synthetic_dataset(n_nodes, n_cells, seed, flip=0.1, relax_steps=4, **kw)
        nodes = nodes_of(specs, with_logic=True)
        parents, luts = compile_network(nodes)
        raw, missing, keyerr = device.ko_ki_sweep(device.BitPlanes.from_packed(planes_np, n_cells), parents, luts)
        w_raw, w_missing, w_key = cport.importance_raw(synth.planes_to_dense(planes_np, n_cells), parents, luts, 50)
        assert int(keyerr.item()) == w_key
        assert (missing.cpu().numpy() == w_missing).all()
        if w_key == 0:
            assert (raw.cpu().numpy() == w_raw).all()


def test_sweep_full_size_linearity_and_shards(cuda):
    """1M cells = a 125,000-cell block repeated 8 times: every output must be exactly 8x the block's,
    and equal to the sum over uneven shards."""
    from scbonita2_b200 import device
    from scbonita2_b200.importance_scores import compile_network
    block_cells = 125_000 - 8          # 124,992 = 3906 words: blocks stay word aligned
    specs, block = synth.synthetic_dataset(N_NODES, block_cells, 5, flip=0.05, relax_steps=30)
    nodes = nodes_of(specs, with_logic=True)
    parents, luts = compile_network(nodes)
    one = device.ko_ki_sweep(device.BitPlanes.from_packed(block, block_cells), parents, luts)
    big_np = np.tile(block, (1, 8))
    big = device.BitPlanes.from_packed(big_np, 8 * block_cells)
    whole = device.ko_ki_sweep(big, parents, luts)
    for a, b in zip(one, whole):
        assert (8 * a == b).all()
    out = None
    n_words = big_np.shape[1]
    for lo, hi in [(0, 4), (4, 20_000), (20_000, n_words)]:
        out = device.ko_ki_sweep(big.word_slice(lo, hi), parents, luts, out=out)
    for a, b in zip(out, whole):
        assert (a == b).all()
    # and a slice small enough for the cell-by-cell C oracle is exact
    sl = 1024
    raw, missing, keyerr = device.ko_ki_sweep(big.word_slice(0, sl // 32), parents, luts)
    w_raw, w_missing, w_key = cport.importance_raw(synth.planes_to_dense(block[:, :sl // 32], sl), parents, luts, 50)
    assert int(keyerr.item()) == w_key and (missing.cpu().numpy() == w_missing).all()
    if w_key == 0:
        assert (raw.cpu().numpy() == w_raw).all()
