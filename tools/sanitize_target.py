"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every
kernel family once, at sizes a sanitizer finishes in a minute or two.

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py

Covers: pack / unpack / chunk majority, the counting kernel (TMA + mbarrier stage ring, dynamic items,
shared-memory atomics, ticket + last-CTA finalize), its sums-only mode + the fitness kernel's in-kernel
inversion, the tag-validated peer protocol with 8 logical ranks (push kernel, last-CTA push, polling
fitness kernel), the importance sweep with all four passes filled (sorted cells, register ring, recheck
in the shared-memory ring, Brent + replay cleanup), trajectories and the pair-distance kernel.  Every
result is checked against the CPU oracle, so a sanitizer run is also a parity run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import cport  # noqa: E402
from tests.helpers import nodes_of, np_counts  # noqa: E402
from tools import synth  # noqa: E402


def main():
    import torch
    from scbonita2_b200 import device
    from scbonita2_b200.distributed import PeerExchange, word_range
    from scbonita2_b200.importance_scores import compile_network
    rng = np.random.default_rng(0)
    # ---- pack / chunk / counts / fitness (both paths) ----
    n, cells = 24, 9_000
    specs, planes_np = synth.synthetic_dataset(n, cells, 3, flip=0.1, relax_steps=2)
    dense = synth.planes_to_dense(planes_np, cells)
    planes = device.BitPlanes.from_dense(dense)
    assert (planes.to_dense() == dense).all()
    perm = rng.permutation(cells)[:8990].astype(np.int32)
    chunked = planes.chunk_majority(perm, 899, 10, 5)
    assert chunked.n_cells == 899
    targets = [s.index for s in specs]
    parents = [synth.spec_parent_rows(s) for s in specs]
    plan = device.RuleCountsPlan(n, targets, parents)
    counts = plan.run(planes).cpu().numpy()
    for g in (0, 7, 23):
        assert (counts[g] == np_counts(dense, targets[g], parents[g])).all()
    lut = torch.from_numpy(rng.integers(0, 256, (40, n), dtype=np.uint8)).cuda()
    want, _ = device.population_fitness(plan.counts, lut)
    for rep in range(2):
        got, _ = plan.fitness_step(planes, lut, chained=(rep == 1))
        assert torch.equal(got, want)
    # ---- peers: 8 logical ranks on one device, both exchange paths ----
    world = 8
    group = PeerExchange.local_group(n, world)
    plans = [device.RuleCountsPlan(n, targets, parents) for _ in range(world)]
    shards = [planes.word_slice(*word_range(cells, r, world)) for r in range(world)]
    for r in range(world):
        plans[r].run(shards[r], exchange=group[r])
    for r in range(world):
        got, _ = group[r].population_fitness(lut)
        assert torch.equal(got, want)
    for r in range(world):
        plans[r].sums_run(shards[r], exchange=group[r])
    for r in range(world):
        got, _ = plans[r].fitness_from_sums(lut, exchange=group[r])
        assert torch.equal(got, want)
    for ex in group:
        ex.check()
        ex.close()
    # ---- importance sweep: a hard case (long cycles, cells without a repeat) so that every pass runs ----
    n2, cells2 = 48, 6_000
    specs2 = synth.random_network(n2, 2, p_inhibit=0.4)
    planes2 = synth.relax_and_flip(specs2, synth.random_planes(n2, cells2, 12), cells2, 2, relax_steps=6, flip=0.05)
    par2, lut2 = compile_network(nodes_of(specs2, with_logic=True))
    raw, missing, key = device.ko_ki_sweep(device.BitPlanes.from_packed(planes2, cells2), par2, lut2, stats=True)
    w_raw, w_missing, w_key = cport.importance_raw_bitsliced(planes2, cells2, par2, lut2, 50)
    assert int(key.item()) == w_key and (missing.cpu().numpy() == w_missing).all() and (raw.cpu().numpy() == w_raw).all()
    print("sweep stats", device.LAST_SWEEP_STATS, "missing normal", int(w_missing[0]), "keyerror", w_key)
    # ---- trajectories + pair distances ----
    traj = device.trajectories(device.BitPlanes.from_packed(planes2, cells2), [0, 5, 77, 5999], par2, lut2, 20)
    assert tuple(traj.shape) == (4, n2, 20)
    pats = device.trajectory_patterns(traj, 2)
    dist = device.pattern_pair_distances(pats, 10)
    assert int(dist[0, 0].item()) == 0
    torch.cuda.synchronize()
    print("sanitize target ok")


if __name__ == "__main__":
    main()
