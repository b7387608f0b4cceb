#!/usr/bin/env python
"""bench.py — GA-fitness throughput and importance-sweep time of the scBONITA2 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], "H1M"): synthetic HIV-dataset-shaped matrix, 200-node
pathway x 1,000,000 cells PER GPU (weak scaling: each rank owns its own cell slice), GA
population 4096.  One step = the fitness of the whole population on all cells:
``scb_rule_counts`` (one pass over the cell bitplanes -> 16 joint counts per node) +, for N > 1,
an NCCL all-reduce of those counts + ``scb_population_fitness``.  metric = cells x nodes x
individuals evaluated per second, whole job.

Also reported on the same JSON line: ``e2e`` (same step through the host-buffer C-ABI:
dense uint8 matrix and truth tables copied from pinned host memory, result copied back),
``roofline`` of the dominant kernel, ``cpu_baseline`` (oracle port on this box's host cores,
bounded sample), and ``sweep`` (importance KO/KI sweep wall time for 200 nodes x 1M cells per GPU).

``--impl reference`` times the reference's own CPU algorithm for the same metric: the oracle
port (eval per cell, oracle/ref_port.py) in a process pool over every host core.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tools import synth  # noqa: E402

METRIC = "ga_fitness_cell_node_individual_evals_per_sec"
UNIT = "evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=1_000_000, help="cells per GPU")
    ap.add_argument("--nodes", type=int, default=200)
    ap.add_argument("--population", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--sweep-steps", type=int, default=2, help="timed sweep repetitions (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl", action="store_true", help="N>1: sum the counts with ncclAllReduce instead of peer memory")
    return ap.parse_args()


def workload_config(args, extra=None):
    cfg = {"workload": "H1M: HIV-shaped synthetic, 200-node pathway x 1M cells per GPU, GA population 4096 "
                       "(BASELINE.json configs[2])",
           "nodes": args.nodes, "cells_per_gpu": args.cells, "population": args.population,
           "sharding": "cells", "seed": args.seed}
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------------------------
def make_nodes(specs):
    from scbonita2_b200.node_class import Node
    nodes = [Node(s.name, s.index, dict(s.predecessors), dict(s.inversions)) for s in specs]
    for node, s in zip(nodes, specs):
        node.best_rule = [node.name, list(node.predecessors), s.logic]
        node.calculation_function = s.logic
    return nodes


def make_population(nodes, population, seed):
    """uint8 [P, N]: every individual picks one of each node's ordered candidate rules."""
    from scbonita2_b200 import rule_compiler
    rng = np.random.default_rng(seed + 17)
    lut = np.empty((population, len(nodes)), dtype=np.uint8)
    for n, node in enumerate(nodes):
        node.find_all_rule_predictions()
        rules = rule_compiler.extend_with_not_variants(node)
        slots = min(3, len(node.predecessors))
        tables = np.array([rule_compiler.compile_for_slots(r[2], slots) for r in rules], dtype=np.uint8)
        lut[:, n] = tables[rng.integers(0, len(tables), population)]
    return lut


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.path = f"/tmp/scb_clocks_{os.getpid()}.csv"
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(gpu_index), "-lms", "100"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            busy = sorted(sm)[len(sm) // 2:]   # upper half: samples taken under load
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on host cores
# ----------------------------------------------------------------------------------------------
_POOL_STATE = {}


def _pool_init(dense, node_args, logics):
    from scbonita2_b200.node_class import Node
    _POOL_STATE["dense"] = dense
    _POOL_STATE["nodes"] = [Node(*a) for a in node_args]
    _POOL_STATE["logics"] = logics


def _pool_task(task):
    from oracle import ref_port
    n, logic = task
    d, c = ref_port.calculate_error(_POOL_STATE["nodes"][n], logic, _POOL_STATE["dense"])
    return d, c


def cpu_fitness_sample(specs, planes_np, n_cells_sample, n_individuals, seed, procs):
    """Time the oracle port (reference algorithm: eval per cell) on a bounded sample:
    ``n_individuals`` individuals x all nodes x ``n_cells_sample`` cells.  Returns evals/s."""
    import multiprocessing as mp
    from scbonita2_b200 import rule_compiler
    dense = synth.planes_to_dense(planes_np[:, :(n_cells_sample + 31) // 32], n_cells_sample).astype(float)
    nodes = make_nodes(specs)
    rng = np.random.default_rng(seed + 5)
    tasks = []
    for _ in range(n_individuals):
        for n, node in enumerate(nodes):
            if node.node_rules is None or not isinstance(node.node_rules[0], list):
                node.find_all_rule_predictions()
                rule_compiler.extend_with_not_variants(node)
            rules = node.node_rules
            tasks.append((n, rules[int(rng.integers(0, len(rules)))][2]))
    node_args = [(s.name, s.index, dict(s.predecessors), dict(s.inversions)) for s in specs]
    t0 = time.perf_counter()
    if procs <= 1:
        _pool_init(dense, node_args, None)
        results = [_pool_task(t) for t in tasks]
    else:
        with mp.Pool(procs, initializer=_pool_init, initargs=(dense, node_args, None)) as pool:
            results = pool.map(_pool_task, tasks, chunksize=max(1, len(tasks) // (procs * 8)))
    dt = time.perf_counter() - t0
    evals = sum(c for _, c in results)
    return evals / dt, dt, len(tasks)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    specs, planes_np = synth.synthetic_dataset(args.nodes, 4096, args.seed, flip=0.05, relax_steps=5)
    sample_cells, sample_ind = 4096, 1
    times, rates = [], []
    for i in range(args.warmup + args.steps):
        rate, dt, n_tasks = cpu_fitness_sample(specs, planes_np, sample_cells, sample_ind, args.seed + i, procs)
        if i >= args.warmup:
            times.append(dt)
            rates.append(rate)
    value = float(np.mean(rates))
    sample = (f"{sample_ind} individual x {args.nodes} nodes x {sample_cells} cells per step "
              f"(oracle port of calculate_error: Python eval per cell), mp.Pool({procs})")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": workload_config(args, {"sample": sample}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from scbonita2_b200 import device, rule_compiler
    from scbonita2_b200.importance_scores import compile_network

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs: same network on every rank, a different cell slice per rank ----
    specs = synth.random_network(args.nodes, args.seed)
    start = synth.random_planes(args.nodes, args.cells, args.seed + 1000 * (rank + 1))
    planes_np = synth.relax_and_flip(specs, start, args.cells, args.seed + rank, relax_steps=30, flip=0.05)
    nodes = make_nodes(specs)
    lut_np = make_population(nodes, args.population, args.seed)
    targets = np.array([n.index for n in nodes], dtype=np.int32)
    parents = np.array([rule_compiler.parent_rows(n) for n in nodes], dtype=np.int32)

    planes = device.BitPlanes.from_packed(planes_np, args.cells)
    d_lut = torch.from_numpy(lut_np).to(dev)
    # The timed steps read their input COLD: the same bitplanes are kept in `n_rot` separate
    # device buffers (together > 2x the 126 MB L2) and step k reads copy k % n_rot, so nothing a
    # step reads is left in L2 by the steps before it.
    plane_bytes = planes.words.numel() * 4
    n_rot = max(2, -(-2 * 126 * 1024 * 1024 // plane_bytes) + 1)
    rot = [planes] + [device.BitPlanes(planes.words.clone(), planes.n_rows, planes.n_cells)
                      for _ in range(n_rot - 1)]
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush_l2():
        flush_buf.fill_(1)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    plan = device.RuleCountsPlan(args.nodes, targets, parents, dev)
    d_diff = torch.empty(args.population, dtype=torch.int64, device=dev)
    # sharded cells: the per-shard counts are summed over peer memory (NVLink), fused into the
    # fitness kernel -- no library collective in the step (--nccl keeps the ncclAllReduce variant)
    exchange = None
    exchange_note = ""
    if world > 1 and not args.nccl:
        from scbonita2_b200.distributed import PeerExchange
        try:
            exchange = PeerExchange.create(args.nodes)
            ok = 1
        except Exception as exc:     # e.g. CUDA IPC not permitted between the processes of this box
            print(f"[bench] rank {rank}: peer-memory exchange unavailable ({exc})", file=sys.stderr)
            ok = 0
        agree = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if int(agree.item()) == 0:   # every rank takes the same path
            if exchange is not None:
                exchange.close()
            exchange = None
            exchange_note = " (peer memory unavailable on this box)"

    def fitness_step(k=0, timers=None):
        # two kernel launches (+ one NCCL all-reduce when sharded); nothing else is enqueued
        if timers is not None:
            timers[0].record()
        counts = plan.run(rot[k % n_rot], exchange=exchange)   # sharded: the last CTA pushes to the peers
        if timers is not None:
            timers[1].record()
        if exchange is not None:
            exchange.population_fitness(d_lut, out=d_diff)
            return d_diff
        if world > 1:
            dist.all_reduce(counts)
        device.population_fitness(counts, d_lut, out=d_diff)
        return d_diff

    units_per_step = float(args.cells) * args.nodes * args.population * world

    for k in range(max(args.warmup, 3)):
        fitness_step(k)
    barrier()
    if exchange is not None:
        # the peer-memory sum against ncclAllReduce of the same counts, once
        exchange.check()
        summed = plan.run(rot[0]).clone()
        dist.all_reduce(summed)
        want, _ = device.population_fitness(summed, d_lut)
        fitness_step(0)
        assert torch.equal(want, d_diff), "peer-memory exchange disagrees with ncclAllReduce"
        barrier()

    def capture(body):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        return g

    # the step is launch-bound (two kernels of a few tens of us): the K steps are ONE CUDA graph
    graph = counts_graph = None
    if (world == 1 or exchange is not None) and not args.no_graph:
        try:
            graph = capture(lambda: [fitness_step(k) for k in range(args.steps)])
            # the dominant kernel alone, K launches, for the roofline line
            counts_graph = capture(lambda: [plan.run(rot[k % n_rot]) for k in range(args.steps)])
        except Exception as exc:  # capture is an optimisation, never a requirement
            print(f"[bench] CUDA graph capture unavailable ({exc}); timing eager launches", file=sys.stderr)
            graph = counts_graph = None
    clocks = ClockSampler(local_rank) if rank == 0 else None

    def timed_region(body):
        """EXACTLY K steps between one pair of CUDA events on the launching stream, bracketed by
        barrier + synchronize; W untimed repetitions of the same region first."""
        for _ in range(2):
            body()
        flush_l2()
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        body()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    wall0 = time.perf_counter()
    eager_total = timed_region(lambda: [fitness_step(k) for k in range(args.steps)])
    eager_ms = eager_total / args.steps
    total_ms = timed_region(graph.replay) if graph is not None else eager_total
    if counts_graph is not None:
        k_ms = timed_region(counts_graph.replay) / args.steps
    else:
        k_ms = timed_region(lambda: [plan.run(rot[k % n_rot]) for k in range(args.steps)]) / args.steps
    # one step on its own (L2 flushed, its own event pair): includes the launch latency that the
    # back-to-back region hides
    iso = []
    for k in range(min(args.steps, 10)):
        flush_l2()
        e0, e1 = ev(), ev()
        e0.record()
        fitness_step(k)
        e1.record()
        iso.append((e0, e1))
    barrier()
    isolated_ms = float(np.median([a.elapsed_time(b) for a, b in iso]))
    wall = time.perf_counter() - wall0
    total_ms = max_over_ranks(float(total_ms))
    ms_per_step = total_ms / args.steps
    value = units_per_step / (ms_per_step * 1e-3)
    checksum = int(d_diff.sum().item())
    del rot[1:]

    # ---- end-to-end through the host-buffer C-ABI (pinned host memory in, result out) ----
    e2e = None
    if not args.no_e2e:
        dense_host = torch.from_numpy(synth.planes_to_dense(planes_np, args.cells)).pin_memory()
        lut_host = torch.from_numpy(lut_np).pin_memory()
        ws = device.HostWorkspace()
        dense_view, lut_view = dense_host.numpy(), lut_host.numpy()

        def e2e_step():
            ws.load_dense(dense_view)                     # H2D + pack kernel
            ws.rule_counts(targets, parents, want=False)  # counts stay on the device
            out, _ = ws.population_fitness(lut_view)      # H2D tables, D2H mismatch totals
            if world > 1:
                t = torch.from_numpy(out).to(dev)
                dist.all_reduce(t)
                out = t.cpu().numpy()
            return out

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            out = e2e_step()
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / n_e2e)
        if world == 1:
            assert int(out.sum()) == checksum, "host-buffer path disagrees with the device path"
        e2e = {"value": units_per_step / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(dense_view.nbytes + lut_view.nbytes + targets.nbytes + parents.nbytes),
               "d2h_bytes_per_step": int(args.population * 8),
               "path": "scb_host_load_dense + scb_host_rule_counts + scb_host_population_fitness"}
        ws.close()
        del dense_host

    # ---- importance sweep (second half of the metric) ----
    sweep = None
    if args.sweep_steps > 0:
        net_parents, net_luts = compile_network(nodes)
        times = []
        for i in range(args.sweep_steps + 1):
            barrier()
            t0 = time.perf_counter()
            out = device.ko_ki_sweep(planes, net_parents, net_luts, steps=50)
            if world > 1:
                for t in out:
                    dist.all_reduce(t)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            if i > 0:
                times.append(dt)
        raw = out[0].cpu().numpy()
        sweep = {"seconds": float(np.mean(times)), "conditions": 2 * args.nodes + 1, "steps": 50,
                 "cells_total": args.cells * world, "raw_checksum": int(raw.sum()),
                 "missing_normal": int(out[1][0].item()),
                 "node_cell_updates_per_s_full_steps": float(
                     (2 * args.nodes + 1) * 50.0 * args.nodes * args.cells * world / np.mean(times))}

    clk = clocks.stop() if clocks else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rule_counts_kernel: one pass over the bitplanes) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    algo_bytes = args.nodes * ((args.cells + 31) // 32) * 4 + args.nodes * 16 * 8
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    traffic_path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(traffic_path) and args.nodes == 200 and args.cells == 1_000_000:
        traffic = json.load(open(traffic_path))["rule_counts_kernel"]["dram_bytes"]   # ncu --set full capture
    roofline = {"kernel": "rule_counts_kernel", "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel_ms": k_ms, "algorithmic_bytes": algo_bytes}

    # integer-pipe view: the per-individual evaluation the reference performs is 4
    # integer instructions per (node, individual, 32 cells); the joint-count statistic replaces it by
    # one pass that does not depend on the population, hence a "fraction" far above 1
    int_peak, int_src = 18081.0, "profiles/r1_int_peaks.jsonl (tools/int_peaks.cu, LOP3 lane-ops on this B200)"
    algo_int = units_per_step / world * 4.0 / 32.0
    roofline_int = {"bound": "int_pipe", "achieved": algo_int / (ms_per_step * 1e-3) / 1e9, "peak": int_peak,
                    "unit": "Gop/s", "frac": algo_int / (ms_per_step * 1e-3) / 1e9 / int_peak,
                    "peak_source": int_src,
                    "note": "algorithmic ops of evaluating every individual's rule on every cell (4 per 32 "
                            "cells); the 16 joint counts per node are a sufficient statistic, so the pass over "
                            "the cells is done once for the whole population"}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        rate, dt, n_tasks = cpu_fitness_sample(specs, planes_np, 12000, 1, args.seed, 1)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"1 individual x {args.nodes} nodes x 12000 cells = {n_tasks} calculate_error "
                                  f"calls (Python eval per cell), {dt:.1f} s, single process like the reference"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args, {
                "exchange": ("none (1 GPU)" if world == 1 else
                             "peer memory: counts pushed over NVLink, summed inside the fitness kernel"
                             if exchange is not None else "ncclAllReduce(int64) between the kernels" + exchange_note),
                "timing": f"K steps back to back between one CUDA event pair; inputs larger than L2: step k reads "
                          f"copy k % {n_rot} of the bitplanes ({n_rot} x {plane_bytes / 1e6:.1f} MB > 2 x 126 MB L2)",
                "wall_s_timed_region": wall, "fitness_checksum": checksum,
                "launch": "one cuda graph of K steps" if graph is not None else "eager",
                "ms_per_step_eager": eager_ms, "ms_per_step_isolated_l2_flushed": isolated_ms}),
            "roofline": roofline, "roofline_int": roofline_int, "cpu_baseline": cpu_baseline, "e2e": e2e, "sweep": sweep,
            "clocks": clk, "gpu_launches": 2 * args.steps}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
