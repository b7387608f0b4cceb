"""GPU suite: the C-ABI fails loudly (negative code + message), never silently."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_error_codes_and_messages(cuda):
    from scbonita2_b200 import _lib, device
    torch = cuda
    L = _lib.lib()
    planes = device.BitPlanes.from_dense((np.arange(3 * 70).reshape(3, 70) % 3 == 0).astype(np.uint8))
    target = torch.tensor([0], dtype=torch.int32, device="cuda")
    parents = torch.tensor([[1, 2, -1]], dtype=torch.int32, device="cuda")
    counts = torch.zeros((1, 16), dtype=torch.int64, device="cuda")
    scratch = torch.zeros(1 << 16, dtype=torch.uint8, device="cuda")
    # scratch too small
    rc = L.scb_rule_counts(planes.words.data_ptr(), 3, planes.wpad, 70, 1, target.data_ptr(), parents.data_ptr(),
                           scratch.data_ptr(), 8, counts.data_ptr(), None)
    assert rc == -4 and b"scratch" in L.scb_last_error()
    # wpad not a multiple of 4
    rc = L.scb_rule_counts(planes.words.data_ptr(), 3, 3, 70, 1, target.data_ptr(), parents.data_ptr(),
                           scratch.data_ptr(), 1 << 16, counts.data_ptr(), None)
    assert rc == -1 and b"wpad" in L.scb_last_error()
    # null pointer
    rc = L.scb_rule_counts(None, 3, planes.wpad, 70, 1, target.data_ptr(), parents.data_ptr(),
                           scratch.data_ptr(), 1 << 16, counts.data_ptr(), None)
    assert rc == -1 and b"null" in L.scb_last_error()
    # sweep: steps out of range, scratch too small
    lut = torch.tensor([0xF0, 0xCC, 0xAA], dtype=torch.uint8, device="cuda")
    par3 = torch.tensor([[0, -1, -1], [0, 1, -1], [2, -1, -1]], dtype=torch.int32, device="cuda")
    raw = torch.zeros(3, dtype=torch.int64, device="cuda")
    miss = torch.zeros(7, dtype=torch.int64, device="cuda")
    key = torch.zeros(1, dtype=torch.int64, device="cuda")
    rc = L.scb_ko_ki_sweep(planes.words.data_ptr(), 3, planes.wpad, 70, par3.data_ptr(), lut.data_ptr(), 99,
                           scratch.data_ptr(), 1 << 16, raw.data_ptr(), miss.data_ptr(), key.data_ptr(), None)
    assert rc == -1 and b"steps" in L.scb_last_error()
    rc = L.scb_ko_ki_sweep(planes.words.data_ptr(), 3, planes.wpad, 70, par3.data_ptr(), lut.data_ptr(), 50,
                           scratch.data_ptr(), 64, raw.data_ptr(), miss.data_ptr(), key.data_ptr(), None)
    assert rc == -4
    # too many nodes for the shared-memory resident sweep: refused, not truncated
    need = L.scb_ko_ki_sweep_scratch_bytes(600, 4, 50)
    big = torch.zeros(int(need), dtype=torch.uint8, device="cuda")
    p600 = torch.zeros((600, 4), dtype=torch.int32, device="cuda")
    par600 = torch.full((600, 3), -1, dtype=torch.int32, device="cuda")
    lut600 = torch.zeros(600, dtype=torch.uint8, device="cuda")
    raw600 = torch.zeros(600, dtype=torch.int64, device="cuda")
    miss600 = torch.zeros(1201, dtype=torch.int64, device="cuda")
    rc = L.scb_ko_ki_sweep(p600.data_ptr(), 600, 4, 100, par600.data_ptr(), lut600.data_ptr(), 50,
                           big.data_ptr(), int(need), raw600.data_ptr(), miss600.data_ptr(), key.data_ptr(), None)
    assert rc == -3 and b"shared-memory" in L.scb_last_error()
    with pytest.raises(_lib.ScbError):
        _lib.check(rc, "scb_ko_ki_sweep")


def test_python_wrappers_validate_indices(cuda):
    from scbonita2_b200 import device
    planes = device.BitPlanes.from_dense(np.ones((4, 40), dtype=np.uint8))
    with pytest.raises(IndexError):
        device.rule_counts(planes, [4], [(0, 1, 2)])
    with pytest.raises(IndexError):
        device.rule_counts(planes, [0], [(0, 9, -1)])
    with pytest.raises(IndexError):
        device.trajectories(planes, [40], np.full((4, 3), -1), np.zeros(4, np.uint8))
    with pytest.raises(ValueError):
        device.ko_ki_sweep(planes, np.full((3, 3), -1), np.zeros(3, np.uint8))
    counts = device.rule_counts(planes, [0], [(1, 2, 3)])
    with pytest.raises(IndexError):
        device.rule_error_table(counts, [1], [0xF0])
    # all-ones rows: every cell reads minterm 7 with target 1
    c = counts.cpu().numpy()[0]
    assert c[15] == 40 and c.sum() == 40


def test_host_workspace_validates(cuda):
    from scbonita2_b200 import _lib, device
    ws = device.HostWorkspace()
    with pytest.raises(_lib.ScbError):
        ws.rule_counts([0], [(0, -1, -1)])                 # nothing loaded yet
    ws.load_dense(np.ones((2, 33), dtype=np.uint8))
    with pytest.raises(_lib.ScbError):
        ws.rule_counts([5], [(0, -1, -1)])                 # row out of range
    ws.rule_counts([0], [(1, -1, -1)])
    with pytest.raises(_lib.ScbError):
        ws.rule_error_table([3], [0xF0])                   # group out of range
    ws.close()
