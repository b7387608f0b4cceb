"""GPU parity: bitplane pack / unpack / majority chunking through the C-ABI."""
import numpy as np
import pytest

from oracle import ref_port
from tests.helpers import golden, matrix_of
from tools import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_rows,n_cells", [(1, 1), (3, 31), (3, 32), (5, 33), (7, 127), (4, 128),
                                            (9, 1000), (2, 4096), (6, 5003), (59, 3621)])
def test_pack_matches_numpy_and_round_trips(cuda, n_rows, n_cells):
    from scbonita2_b200.device import BitPlanes
    rng = np.random.default_rng(n_rows * 1000 + n_cells)
    dense = (rng.random((n_rows, n_cells)) < 0.37).astype(np.uint8)
    planes = BitPlanes.from_dense(dense)
    assert planes.wpad % 4 == 0 and planes.wpad * 32 >= n_cells
    want = synth.dense_to_planes(dense)
    got = planes.words.cpu().numpy().view(np.uint32)
    assert (got[:, :want.shape[1]] == want).all()
    assert (got[:, want.shape[1]:] == 0).all()          # padding words are zero
    assert (planes.to_dense() == dense).all()
    assert (planes.to_packed() == want).all()
    # any non-zero byte counts as 1; float / bool / sparse inputs pack identically
    import scipy.sparse as sp
    assert (BitPlanes.from_dense(dense * 7).to_packed() == want).all()
    assert (BitPlanes.from_dense(sp.csr_matrix(dense.astype(float))).to_packed() == want).all()
    assert (BitPlanes.from_packed(want, n_cells).to_dense() == dense).all()


def test_pack_unaligned_source_uses_generic_path(cuda):
    """Row stride not a multiple of 16 bytes -> byte-wise kernel; must agree with the fast one."""
    from scbonita2_b200 import _lib
    from scbonita2_b200.device import padded_words
    torch = cuda
    rng = np.random.default_rng(5)
    for n_cells in (45, 1001, 2049):
        dense = (rng.random((6, n_cells)) < 0.5).astype(np.uint8)
        d = torch.from_numpy(dense).cuda()
        wpad = padded_words(n_cells)
        out = torch.full((6, wpad), -1, dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().scb_pack_bitplanes(d.data_ptr(), 6, n_cells, n_cells, out.data_ptr(),
                                                 wpad, torch.cuda.current_stream().cuda_stream))
        want = synth.dense_to_planes(dense)
        got = out.cpu().numpy().view(np.uint32)
        assert (got[:, :want.shape[1]] == want).all() and (got[:, want.shape[1]:] == 0).all()


def test_pack_rejects_bad_arguments(cuda):
    from scbonita2_b200 import _lib
    torch = cuda
    d = torch.zeros((2, 64), dtype=torch.uint8, device="cuda")
    out = torch.zeros((2, 4), dtype=torch.int32, device="cuda")
    rc = _lib.lib().scb_pack_bitplanes(d.data_ptr(), 2, 64, 64, out.data_ptr(), 3, None)
    assert rc == -1 and b"wpad" in _lib.lib().scb_last_error()
    with pytest.raises(_lib.ScbError):
        _lib.check(_lib.lib().scb_pack_bitplanes(None, 2, 64, 64, out.data_ptr(), 4, None))


def test_chunk_majority_golden_widths(cuda):
    from scbonita2_b200.device import BitPlanes
    from scbonita2_b200.rule_determination import chunk_layout, chunk_min_ones
    for case in golden("chunk_widths.json"):
        dense = matrix_of(case["matrix"])
        perm, n_chunks, size = chunk_layout(case["n_cells"], case["num_chunks"])
        out = BitPlanes.from_dense(dense).chunk_majority(perm, n_chunks, size, chunk_min_ones(size))
        assert out.n_cells == n_chunks
        assert (out.to_dense() == matrix_of(case["chunked"])).all(), case["n_cells"]


def test_chunk_majority_random_vs_oracle(cuda):
    from scbonita2_b200.device import BitPlanes
    from scbonita2_b200.rule_determination import chunk_layout, chunk_min_ones
    rng = np.random.default_rng(9)
    for n_cells, num_chunks in [(2500, 250), (3621, 362), (10007, 1001), (999, 999)]:
        dense = (rng.random((4, n_cells)) < 0.45).astype(np.uint8)
        perm, n_chunks, size = chunk_layout(n_cells, num_chunks)
        out = BitPlanes.from_dense(dense).chunk_majority(perm, n_chunks, size, chunk_min_ones(size))
        assert (out.to_dense() == ref_port.chunk_data(dense, num_chunks).astype(np.uint8)).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,threshold", [((1, 1), 0.0), ((3, 31), 0.5), ((5, 1000), 0.01), ((59, 3621), 0.001),
                                             ((7, 40_003), 0.3)])
def test_binarize_pack_matches_sklearn_recipe(cuda, dtype, shape, threshold):
    """scb_binarize_pack_f32/f64 == csr + binarize(threshold) of the reference (value > threshold)."""
    from scbonita2_b200 import device
    rng = np.random.default_rng(shape[1])
    x = rng.normal(0.2, 1.0, shape).astype(dtype)
    x[rng.random(shape) < 0.3] = 0
    x.flat[::97] = threshold            # values AT the threshold stay 0
    if x.size > 5:
        x.flat[3] = np.nan
    planes = device.BitPlanes.from_float(x, threshold)
    want = (x > dtype(threshold)).astype(np.uint8)
    assert (planes.to_dense() == want).all()
    tail = planes.words.cpu().numpy().view(np.uint32)
    n_words = (shape[1] + 31) // 32
    if shape[1] % 32:
        assert (tail[:, n_words - 1] >> (shape[1] % 32) == 0).all()
    assert (tail[:, n_words:] == 0).all()
