"""Fitness kernel with sharded cells on ONE device (8 logical ranks wired directly): time of the polling
fitness kernel alone, per CTA-count setting (SCB_FITNESS_PER_WARP), for both exchange formats."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import synth
from scbonita2_b200 import device
from scbonita2_b200.distributed import PeerExchange, word_range

n, cells, world, K = 200, 200_000, int(os.environ.get("WORLD", 8)), 20
specs = synth.random_network(n, 0)
planes = device.BitPlanes.from_packed(synth.random_planes(n, cells, 1000), cells)
targets, parents = [s.index for s in specs], [synth.spec_parent_rows(s) for s in specs]
group = PeerExchange.local_group(n, world)
plans = [device.RuleCountsPlan(n, targets, parents) for _ in range(world)]
shards = [planes.word_slice(*word_range(cells, r, world)) for r in range(world)]
lut = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (4096, n), dtype=np.uint8)).cuda()
out = torch.empty(4096, dtype=torch.int64, device="cuda")
for r in range(world):
    plans[r].run(shards[r], exchange=group[r])
torch.cuda.synchronize()
def timed(body):
    for _ in range(3): body()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): body()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K
t_counts = timed(lambda: group[0].population_fitness(lut, out=out))
for r in range(world):
    plans[r].sums_run(shards[r], exchange=group[r])
torch.cuda.synchronize()
t_sums = timed(lambda: plans[0].fitness_from_sums(lut, out=out, exchange=group[0]))
print(f"world={world} per_warp={os.environ.get('SCB_FITNESS_PER_WARP','default')}: fitness from peers' counts {t_counts:.1f} us, from peers' sums {t_sums:.1f} us (eager, back to back)")
