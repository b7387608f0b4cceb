"""Phase timeline of rule_counts_kernel (needs a library built with SCB_TRACE=1):
per-CTA clock64 stamps -> median phase lengths in cycles."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import synth
from scbonita2_b200 import device, _lib

def main():
    n = int(os.environ.get("N_NODES", 200)); cells = int(os.environ.get("N_CELLS", 1_000_000))
    specs = synth.random_network(n, 0)
    planes = device.BitPlanes.from_packed(synth.random_planes(n, cells, 1000), cells)
    plan = device.RuleCountsPlan(n, [s.index for s in specs], [synth.spec_parent_rows(s) for s in specs])
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        plan.run(planes)
    flush.fill_(1)
    plan.run(planes)
    torch.cuda.synchronize()
    h = ctypes.CDLL(_lib.LIB_PATH)
    buf = np.zeros((256, 12), dtype=np.int64)
    h.scb_debug_trace_read(buf.ctypes.data_as(ctypes.c_void_p))
    n_cta = min(148, (planes.wpad + 31) // 32)
    t = buf[:n_cta].astype(np.float64)
    names = ["entry->init sync", "->setup done", "->first tile", "->loop exit", "->cta sync", "->flush+fence", "->ticket"]
    for k, name in enumerate(names):
        d = t[:, k + 1] - t[:, k]
        print(f"{name:18s} cycles: median {np.median(d):9.0f}  min {d.min():9.0f}  max {d.max():9.0f}")
    last = np.argmax(t[:, 9] > 0) if (t[:, 9] > 0).any() else -1
    lasts = np.nonzero(buf[:n_cta, 8] > buf[:n_cta, 7])[0]
    for c in lasts[-1:]:
        print(f"last CTA {c}: finalize {buf[c,8]-buf[c,7]} cycles (loads issued at +{buf[c,10]-buf[c,7]}, landed at +{buf[c,11]-buf[c,7]}), clean {buf[c,9]-buf[c,8]} cycles")
    print(f"total entry->ticket: median {np.median(t[:,7]-t[:,0]):.0f} max {(t[:,7]-t[:,0]).max():.0f} cycles")

main()
