#!/usr/bin/env python
"""bench.py -- GA-fitness throughput and importance-sweep time of the scBONITA2 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], "H1M"): synthetic HIV-dataset-shaped matrix, 200-node
pathway x 1,000,000 cells PER GPU (weak scaling: each rank owns its own cell slice), GA
population 4096.  One step = the fitness of the whole population on all cells:
``scb_rule_counts`` (one pass over the cell bitplanes -> 16 joint counts per node; sharded: the
counts are pushed to the peers over NVLink by the kernel's last CTA) + ``scb_population_fitness``.
metric = cells x nodes x individuals evaluated per second, whole job.

On the same JSON line:
  e2e           the same step through the host-buffer C-ABI (``scb_host_*``): packed bitplanes and truth
                tables copied from pinned host memory, mismatch totals copied back (``e2e_dense``: the same
                from the dense uint8 matrix, 8 bytes moved per bit of information)
  roofline      the dominant kernel (rule_counts_kernel) against the roofline that binds it: HBM bytes
                or required integer-pipe instructions of the subset-sum formulation (``roofline_hbm``,
                ``roofline_pipe`` carry both)
  cpu_baseline  the reference's own ``calculate_error`` (baseline/_ref, else the oracle port) on host cores
  sweep         second half of the metric: the 2N+1-condition importance sweep (BASELINE.json configs[3],
                imp-1M) with its own roofline / cpu_baseline / e2e
  strong        N > 1: the north-star strong-scaling view -- the SAME 1M cells split over the N ranks
                (GA step and sweep), beside the weak-scaling headline

``--impl reference`` times the reference's own CPU implementation of the path (the unmodified
sources installed in baseline/_ref by ``__graft_entry__.build()``; the oracle port when absent) on a
bounded sample with every host core.
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
    ap.add_argument("--fused-peers", action="store_true", help="N>1: exchange subset sums (scb_rule_sums_run) instead of counts")
    ap.add_argument("--unfused", action="store_true", help="GA step = scb_rule_counts_run + scb_population_fitness (round-1 kernels)")
    ap.add_argument("--no-e2e-dense", action="store_true", help="skip the dense-uint8 variant of the end-to-end step")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the strong-scaling block (same 1M cells split over the ranks)")
    ap.add_argument("--cpu-leg", default="", choices=["", "fitness", "sweep"], help=argparse.SUPPRESS)
    ap.add_argument("--leg-cells", type=int, default=4096, help=argparse.SUPPRESS)
    ap.add_argument("--leg-procs", type=int, default=1, help=argparse.SUPPRESS)
    ap.add_argument("--leg-rep", type=int, default=0, help=argparse.SUPPRESS)
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
# CPU legs (reference arm, cpu_baseline): the reference's own code on host cores.  They run in a
# child process (``bench.py --cpu-leg ...``): the reference forks a process pool, which must not
# happen in a process that holds a CUDA context and NCCL threads.
# ----------------------------------------------------------------------------------------------
_POOL_STATE = {}


def reference_modules():
    """(namespace, "reference") with the UNMODIFIED reference modules (baseline/_ref or /root/reference,
    through the oracle/ref_loader shims), or (None, "port") when neither exists."""
    from oracle import ref_loader
    if ref_loader.available():
        try:
            return ref_loader.load(), "reference"
        except Exception as exc:  # pragma: no cover - depends on the box
            print(f"[bench] reference import failed ({exc}); timing the oracle port", file=sys.stderr)
    return None, "port"


def _pool_init(dense, node_args, kind):
    import logging
    logging.disable(logging.CRITICAL)
    _POOL_STATE["dense"] = dense
    if kind == "reference":
        ns, _ = reference_modules()
        _POOL_STATE["nodes"] = [ns.node_class.Node(*a) for a in node_args]
        _POOL_STATE["error"] = ns.rule_determination.RuleDetermination.calculate_error
    else:
        from oracle import ref_port
        from scbonita2_b200.node_class import Node
        _POOL_STATE["nodes"] = [Node(*a) for a in node_args]
        _POOL_STATE["error"] = ref_port.calculate_error


def _pool_ready(_):
    time.sleep(0.05)
    return 0


def _pool_task(task):
    n, logic = task
    d, c = _POOL_STATE["error"](_POOL_STATE["nodes"][n], logic, _POOL_STATE["dense"])
    return int(d), int(c)


def cpu_fitness_sample(specs, planes_np, n_cells_sample, n_individuals, seed, procs, kind):
    """Time ``RuleDetermination.calculate_error`` (rule_determination.py:509-582: Python ``eval`` per
    cell) on a bounded sample: ``n_individuals`` individuals x all nodes x ``n_cells_sample`` cells.
    Returns (evals/s, seconds, calls)."""
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
    if procs <= 1:
        _pool_init(dense, node_args, kind)
        t0 = time.perf_counter()
        results = [_pool_task(t) for t in tasks]
        dt = time.perf_counter() - t0
    else:
        with mp.Pool(procs, initializer=_pool_init, initargs=(dense, node_args, kind)) as pool:
            pool.map(_pool_ready, range(4 * procs))     # workers up and the reference imported: not timed
            t0 = time.perf_counter()
            results = pool.map(_pool_task, tasks, chunksize=max(1, len(tasks) // (procs * 8)))
            dt = time.perf_counter() - t0
    evals = sum(c for _, c in results)
    return evals / dt, dt, len(tasks)


def cpu_sweep_sample(specs, planes_np, n_cells_sample, tag, kind, steps=50):
    """Time the importance sweep on ``n_cells_sample`` cells with the reference's own
    ``CalculateImportanceScore.calculate_importance_scores()`` (importance_scores.py:206-253,256-323:
    2N+1 simulations under ``mp.Pool(mp.cpu_count())``, pickles, Python scoring), or with the oracle's
    64-cells-per-word C port on every core.  Returns (seconds, raw totals, mean aligned length)."""
    import logging
    import random
    from oracle import cport
    from scbonita2_b200.importance_scores import compile_network
    nodes = make_nodes(specs)
    parents, luts = compile_network(nodes)
    words = planes_np[:, :(n_cells_sample + 31) // 32].copy()
    synth.mask_tail(words, n_cells_sample)
    _, _, _, (lam_sum, lam_cnt) = cport.importance_raw_bitsliced(words, n_cells_sample, parents, luts, steps, with_stats=True)
    mean_len = lam_sum / max(1, lam_cnt)
    if kind == "reference":
        import scipy.sparse as sp
        logging.disable(logging.CRITICAL)
        ns, _ = reference_modules()
        dense = synth.planes_to_dense(words, n_cells_sample)
        ref_nodes = [ns.node_class.Node(s.name, s.index, dict(s.predecessors), dict(s.inversions)) for s in specs]
        for node, s in zip(ref_nodes, specs):
            node.best_rule = [node.name, [p for p in node.predecessors], s.logic]
        random.seed(0)
        calc = ns.importance_scores.CalculateImportanceScore(ref_nodes, sp.csr_matrix(dense.astype(float)).astype(bool),
                                                             f"net_{tag}", f"ds_{tag}")   # fresh names: no cached pickles
        t0 = time.perf_counter()
        calc.calculate_importance_scores()
        return time.perf_counter() - t0, len(calc.cell_sample_indices), mean_len
    t0 = time.perf_counter()
    cport.importance_raw_bitsliced(words, n_cells_sample, parents, luts, steps)
    return time.perf_counter() - t0, n_cells_sample, mean_len


def run_cpu_leg(args):
    """Child process of the bench: prints one JSON object for the requested CPU leg."""
    _, kind = reference_modules()
    procs = os.cpu_count() or 1
    specs, planes_np = synth.synthetic_dataset(args.nodes, max(args.leg_cells, 32), args.seed, flip=0.05, relax_steps=30)
    if args.cpu_leg == "fitness":
        rate, dt, n_tasks = cpu_fitness_sample(specs, planes_np, args.leg_cells, 1, args.seed + args.leg_rep,
                                               args.leg_procs, kind)
        out = {"kind": kind, "value": rate, "seconds": dt, "calls": n_tasks, "cores": args.leg_procs,
               "cells": args.leg_cells}
    else:
        dt, used, mean_len = cpu_sweep_sample(specs, planes_np, args.leg_cells, f"bench{os.getpid()}_{args.leg_rep}", kind)
        out = {"kind": kind, "seconds": dt, "cells": int(used), "cores": procs, "mean_aligned_length": mean_len}
    print("CPU_LEG " + json.dumps(out), flush=True)


def cpu_leg(leg, args, cells, procs=1, rep=0, timeout=900):
    """Run one CPU leg in a child process and return its JSON (None if it failed)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-leg", leg, "--leg-cells", str(cells),
           "--leg-procs", str(procs), "--leg-rep", str(rep), "--nodes", str(args.nodes), "--seed", str(args.seed)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "CUDA_VISIBLE_DEVICES"):
        env.pop(k, None)
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        for line in proc.stdout.splitlines():
            if line.startswith("CPU_LEG "):
                return json.loads(line[8:])
        print(f"[bench] cpu leg {leg} produced no result: {proc.stderr[-400:]}", file=sys.stderr)
    except Exception as exc:
        print(f"[bench] cpu leg {leg} failed: {exc}", file=sys.stderr)
    return None


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path on this box's host
    cores, same metric and config, every step a bounded sample (scale stated in ``sample``)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    sample_cells = 4096
    times, rates, kind = [], [], "port"
    for i in range(args.warmup + args.steps):
        leg = cpu_leg("fitness", args, sample_cells, procs=procs, rep=i)
        if leg is None:
            continue
        kind = leg["kind"]
        if i >= args.warmup:
            times.append(leg["seconds"])
            rates.append(leg["value"])
    value = float(np.mean(rates)) if rates else 0.0
    what = ("unmodified reference RuleDetermination.calculate_error (baseline/_ref)" if kind == "reference"
            else "oracle port of calculate_error")
    sample = (f"1 individual x {args.nodes} nodes x {sample_cells} cells per step ({what}: Python eval per cell), "
              f"mp.Pool({procs}); full step = {args.population} individuals x {args.cells} cells: "
              f"scale factor {args.population * args.cells / sample_cells:.0f}")
    sweep = None
    if args.sweep_steps > 0:
        leg = cpu_leg("sweep", args, 500, rep=0)
        if leg is not None:
            sweep = {"seconds_sample": leg["seconds"], "cells_sample": leg["cells"], "kind": leg["kind"],
                     "cores": leg["cores"], "scale_factor": args.cells / leg["cells"],
                     "seconds_extrapolated": leg["seconds"] * args.cells / leg["cells"],
                     "what": "CalculateImportanceScore.calculate_importance_scores() under its own mp.Pool"
                             if leg["kind"] == "reference" else "oracle C port (64 cells per word), all cores"}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3) if times else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": workload_config(args, {"sample": sample}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "sweep": sweep, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# roofline bookkeeping
# ----------------------------------------------------------------------------------------------
def measured_peaks():
    """HBM GB/s from MEASURED_PEAKS.json (driver-written) and the integer-pipe rates measured on a B200
    of this pool with tools/int_peaks.cu (lane-ops per second; a lane-op is one 32-bit word = 32 cells)."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        hbm, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    pipes, src = {"lop3": 18081.0, "popc": 4622.0}, "profiles/r1_int_peaks.jsonl"
    for name in ("r2_int_peaks.jsonl", "r1_int_peaks.jsonl"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            for line in open(path):
                try:
                    rec = json.loads(line)
                except ValueError:
                    continue
                if rec.get("pipe") in ("lop3", "popc", "imad"):
                    pipes[rec["pipe"]] = float(rec["gops"])
            src = f"profiles/{name}"
            break
    return hbm, hbm_src, pipes, src + " (tools/int_peaks.cu on this pool's B200: lane-ops/s, every SM busy)"


def counts_required_ops(parents, n_rows, n_words):
    """Lane-ops the subset-sum formulation of rule_counts_kernel REQUIRES: per AND
    term and 8 words 8 AND + 4 carry-save adders (8 LOP3) on the ALU pipe and 4 POPC on the XU pipe; the
    row popcounts cost the adders and the POPCs without the ANDs.  Terms per group: 1 / 4 / 11 for
    1 / 2 / 3 bound parents."""
    arity = (np.asarray(parents) >= 0).sum(axis=1)
    terms = int(sum({0: 0, 1: 1, 2: 4, 3: 11}[int(a)] for a in arity))
    alu = (2.0 * terms + 1.0 * n_rows) * n_words
    xu = 0.5 * (terms + n_rows) * n_words
    return alu, xu, terms


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from scbonita2_b200 import device, distributed as scb_dist, rule_compiler
    from scbonita2_b200.importance_scores import compile_network

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    numa_bound = device.bind_host_to_device()   # pinned host buffers on the GPU's own NUMA node
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

    # ---- inputs: same network on every rank; weak scaling: a different 1M-cell slice per rank ----
    specs = synth.random_network(args.nodes, args.seed)

    def make_cells(seed_rank):
        start = synth.random_planes(args.nodes, args.cells, args.seed + 1000 * (seed_rank + 1))
        return synth.relax_and_flip(specs, start, args.cells, args.seed + seed_rank, relax_steps=30, flip=0.05)

    planes_np = make_cells(rank)
    nodes = make_nodes(specs)
    lut_np = make_population(nodes, args.population, args.seed)
    targets = np.array([n.index for n in nodes], dtype=np.int32)
    parents = np.array([rule_compiler.parent_rows(n) for n in nodes], dtype=np.int32)
    net_parents, net_luts = compile_network(nodes)
    d_lut = torch.from_numpy(lut_np).to(dev)
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush_l2():
        flush_buf.fill_(1)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

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

    plan = device.RuleCountsPlan(args.nodes, targets, parents, dev)
    d_diff = torch.empty(args.population, dtype=torch.int64, device=dev)

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

    def timed_region(body):
        """EXACTLY K steps between one pair of CUDA events on the launching stream, bracketed by
        barrier + synchronize; untimed repetitions of the same region first."""
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

    def ga_bench(planes_host, n_cells_local):
        """The GA fitness step on one cell slice per rank.  The timed steps read their input COLD: the
        same bitplanes are kept in n_rot device buffers (together > 2x the 126 MB L2) and step k reads
        copy k % n_rot.  Returns a dict of timings (ms) and the checksum."""
        planes = device.BitPlanes.from_packed(planes_host, n_cells_local)
        plane_bytes = planes.words.numel() * 4
        n_rot = max(2, -(-2 * 126 * 1024 * 1024 // plane_bytes) + 1)
        rot = [planes] + [device.BitPlanes(planes.words.clone(), planes.n_rows, planes.n_cells)
                          for _ in range(n_rot - 1)]

        # one GPU: counting kernel in sums-only mode + fitness kernel that inverts (no ticket / last-CTA finalize).
        # Sharded: the last CTA is needed for the push either way, and the counts format (two words per
        # thread in the polling fitness kernel) measured faster at N=8 (28.9 vs 34.5 us per step): round-1 pair.
        fused = plan.fused_supported and not args.unfused and (world == 1 or args.fused_peers)

        def fitness_step(k=0):
            # two kernel launches (+ one NCCL all-reduce in the --nccl variant); nothing else is enqueued
            if fused:   # sums kernel (sharded: its last CTA pushes to the peers) + fitness kernel that inverts them
                plan.fitness_step(rot[k % n_rot], d_lut, out=d_diff, exchange=exchange, chained=True)
                return d_diff
            counts = plan.run(rot[k % n_rot], exchange=exchange, chained=True)
            if exchange is not None:
                exchange.population_fitness(d_lut, out=d_diff)
                return d_diff
            if world > 1:
                dist.all_reduce(counts)
            device.population_fitness(counts, d_lut, out=d_diff)
            return d_diff

        def counts_only(k):
            # the dominant kernel alone (its consumer is missing: the sums just pile up, the timing is the same)
            if fused:
                plan.sums_run(rot[k % n_rot], chained=True)
            else:
                plan.run(rot[k % n_rot], chained=True)

        for k in range(max(args.warmup, 3)):
            fitness_step(k)
        barrier()
        if exchange is not None:
            # the peer-memory sum against ncclAllReduce of the same counts, once
            exchange.check()
            torch.cuda.synchronize()
            summed = plan.run(rot[0]).clone()
            dist.all_reduce(summed)
            want, _ = device.population_fitness(summed, d_lut)
            fitness_step(0)
            assert torch.equal(want, d_diff), "peer-memory exchange disagrees with ncclAllReduce"
            barrier()
        # the step is launch-bound (two kernels of a few tens of us): the K steps are ONE CUDA graph
        graph = counts_graph = None
        if (world == 1 or exchange is not None) and not args.no_graph:
            try:
                graph = capture(lambda: [fitness_step(k) for k in range(args.steps)])
                # the dominant kernel alone, K launches, for the roofline line
                counts_graph = capture(lambda: [counts_only(k) for k in range(args.steps)])
                plan.scratch.zero_()
            except Exception as exc:  # capture is an optimisation, never a requirement
                print(f"[bench] CUDA graph capture unavailable ({exc}); timing eager launches", file=sys.stderr)
                graph = counts_graph = None
        wall0 = time.perf_counter()
        eager_total = timed_region(lambda: [fitness_step(k) for k in range(args.steps)])
        total_ms = timed_region(graph.replay) if graph is not None else eager_total
        if counts_graph is not None:
            k_ms = timed_region(counts_graph.replay) / args.steps
        else:
            k_ms = timed_region(lambda: [counts_only(k) for k in range(args.steps)]) / args.steps
        plan.scratch.zero_()   # (the sums kernel timed alone leaves its sums behind)
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
        if exchange is not None:
            exchange.check()     # a rank that gave up waiting poisons its results AND raises here
        out = {"ms_per_step": max_over_ranks(float(total_ms)) / args.steps, "eager_ms": eager_total / args.steps,
               "counts_kernel_ms": k_ms, "isolated_ms": float(np.median([a.elapsed_time(b) for a, b in iso])),
               "wall": time.perf_counter() - wall0, "checksum": int(d_diff.sum().item()), "n_rot": n_rot,
               "plane_bytes": plane_bytes, "graph": graph is not None, "planes": planes,
               "kernel": "rule_sums_kernel" if fused else "rule_counts_kernel"}
        del rot[1:]
        return out

    def sweep_bench(planes, reps):
        """The 2N+1-condition importance sweep on this rank's slice + the integer all-reduce."""
        times, out = [], None
        for i in range(reps + 1):
            barrier()
            e0, e1 = ev(), ev()
            e0.record()
            out = device.ko_ki_sweep(planes, net_parents, net_luts, steps=50)
            if world > 1:
                scb_dist.all_reduce_sum(out)
            e1.record()
            torch.cuda.synchronize()
            dt = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
            if i > 0:
                times.append(dt)
        return float(np.mean(times)), out

    clocks = ClockSampler(local_rank) if rank == 0 else None

    # ---- headline: GA fitness step, weak scaling (1M cells per GPU) ----
    ga = ga_bench(planes_np, args.cells)
    planes = ga.pop("planes")
    units_per_step = float(args.cells) * args.nodes * args.population * world
    ms_per_step = ga["ms_per_step"]
    value = units_per_step / (ms_per_step * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI (pinned host memory in, result out) ----
    e2e = e2e_dense = None
    n_words = (args.cells + 31) // 32
    if not args.no_e2e:
        packed_host = torch.from_numpy(np.ascontiguousarray(planes_np[:, :n_words]).view(np.int32)).pin_memory()
        lut_host = torch.from_numpy(lut_np).pin_memory()
        ws = device.HostWorkspace()
        packed_view, lut_view = packed_host.numpy().view(np.uint32), lut_host.numpy()

        def e2e_step(load):
            load()                                        # H2D (+ pack kernel for the dense matrix)
            ws.rule_counts(targets, parents, want=False)  # counts stay on the device
            out, _ = ws.population_fitness(lut_view)      # H2D tables, D2H mismatch totals
            if world > 1:
                t = torch.from_numpy(out).to(dev)
                dist.all_reduce(t)
                out = t.cpu().numpy()
            return out

        def time_e2e(load, in_bytes, path):
            for _ in range(max(1, min(args.warmup, 2))):
                e2e_step(load)
            barrier()
            n_e2e = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                out = e2e_step(load)
            barrier()
            e2e_s = max_over_ranks((time.perf_counter() - t0) / n_e2e)
            if world == 1:
                assert int(out.sum()) == ga["checksum"], "host-buffer path disagrees with the device path"
            return {"value": units_per_step / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(in_bytes + lut_view.nbytes + targets.nbytes + parents.nbytes),
                    "d2h_bytes_per_step": int(args.population * 8), "path": path,
                    "host_numa_bound_to_gpu": bool(numa_bound)}

        e2e = time_e2e(lambda: ws.load_planes(packed_view, args.cells), packed_view.nbytes,
                       "scb_host_load_planes (packed uint32 cell bitplanes, the layout the Python mirrors keep) + "
                       "scb_host_rule_counts + scb_host_population_fitness")
        if not args.no_e2e_dense:
            dense_host = torch.from_numpy(synth.planes_to_dense(planes_np, args.cells)).pin_memory()
            dense_view = dense_host.numpy()
            e2e_dense = time_e2e(lambda: ws.load_dense(dense_view), dense_view.nbytes,
                                 "scb_host_load_dense (dense uint8 0/1 matrix: 8 bytes per bit) + "
                                 "scb_host_rule_counts + scb_host_population_fitness")
            del dense_host

    # ---- importance sweep (second half of the metric), weak: 1M cells per GPU ----
    sweep = None
    if args.sweep_steps > 0:
        sweep_s, out = sweep_bench(planes, args.sweep_steps)
        raw = out[0].cpu().numpy()
        sweep = {"seconds": sweep_s, "conditions": 2 * args.nodes + 1, "steps": 50,
                 "cells_total": args.cells * world, "raw_checksum": int(raw.sum()),
                 "missing_normal": int(out[1][0].item()),
                 "timing": "CUDA events around scb_ko_ki_sweep (+ the integer all-reduce), max over ranks"}
        if not args.no_e2e:
            # end to end: packed planes from pinned host memory in, integer totals out, every repetition
            def sweep_e2e_step():
                ws.load_planes(packed_view, args.cells)
                r, m, k = ws.ko_ki_sweep(net_parents, net_luts, 50)
                if world > 1:
                    t = torch.from_numpy(np.concatenate([r, m, k])).to(dev)
                    dist.all_reduce(t)
                    r = t.cpu().numpy()[:args.nodes]
                return r
            sweep_e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(max(1, args.sweep_steps)):
                r = sweep_e2e_step()
            barrier()
            e2e_sweep_s = max_over_ranks((time.perf_counter() - t0) / max(1, args.sweep_steps))
            assert int(np.asarray(r).sum()) == sweep["raw_checksum"], "host-buffer sweep disagrees with the device path"
            sweep["e2e"] = {"seconds": e2e_sweep_s, "h2d_bytes": int(packed_view.nbytes + net_parents.nbytes + net_luts.nbytes),
                            "d2h_bytes": int((3 * args.nodes + 2) * 8),
                            "path": "scb_host_load_planes + scb_host_ko_ki_sweep (host buffers in, totals out)"}
    if not args.no_e2e:
        ws.close()

    # ---- strong scaling (north star): the SAME 1M cells split over the ranks ----
    strong = None
    if world > 1 and not args.no_strong:
        global_np = make_cells(0) if rank != 0 else planes_np     # rank 0's weak slice IS the N=1 data set
        local_np, local_cells = scb_dist.shard_packed(global_np[:, :n_words], args.cells, rank, world)
        del planes
        sga = ga_bench(local_np, local_cells)
        splanes = sga.pop("planes")
        strong = {"cells_total": args.cells, "cells_per_gpu": local_cells, "ga_ms_per_step": sga["ms_per_step"],
                  "ga_value": float(args.cells) * args.nodes * args.population / (sga["ms_per_step"] * 1e-3),
                  "ga_counts_kernel_ms": sga["counts_kernel_ms"], "fitness_checksum": sga["checksum"],
                  "note": "the GA step is two kernels of a few microseconds of fixed cost each: it cannot "
                          "strong-scale at 1M cells; the sweep does"}
        if args.sweep_steps > 0:
            s_s, s_out = sweep_bench(splanes, args.sweep_steps)
            strong.update(sweep_seconds=s_s, sweep_raw_checksum=int(s_out[0].sum().item()))

    clk = clocks.stop() if clocks else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rule_counts_kernel: one pass over the bitplanes) ----
    hbm_peak, hbm_src, pipes, pipe_src = measured_peaks()
    k_ms = ga["counts_kernel_ms"]
    algo_bytes = args.nodes * n_words * 4 + args.nodes * 16 * 8
    traffic = None
    for name in ("r2_traffic.json", "r1_traffic.json"):       # ncu --set full capture of one launch on H1M
        traffic_path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(traffic_path) and args.nodes == 200 and args.cells == 1_000_000:
            traffic = json.load(open(traffic_path))["rule_counts_kernel"]["dram_bytes"]
            break
    roofline_hbm = {"kernel": ga["kernel"], "bound": "hbm", "achieved": algo_bytes / (k_ms * 1e-3) / 1e9,
                    "peak": hbm_peak, "unit": "GB/s", "frac": algo_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                    "traffic": traffic, "peak_source": hbm_src, "kernel_ms": k_ms, "algorithmic_bytes": algo_bytes}
    alu_ops, xu_ops, n_terms = counts_required_ops(parents, args.nodes, n_words)
    t_hbm = algo_bytes / (hbm_peak * 1e9)
    t_alu = alu_ops / (pipes["lop3"] * 1e9)
    t_xu = xu_ops / (pipes["popc"] * 1e9)
    binding = max((t_hbm, "hbm"), (t_alu, "alu"), (t_xu, "xu"))
    pipe_name, pipe_ops, pipe_peak = ("xu", xu_ops, pipes["popc"]) if t_xu >= t_alu else ("alu", alu_ops, pipes["lop3"])
    roofline_pipe = {"kernel": ga["kernel"], "bound": "int_pipe_" + pipe_name,
                     "achieved": pipe_ops / (k_ms * 1e-3) / 1e9, "peak": pipe_peak, "unit": "Glane-op/s",
                     "frac": pipe_ops / (k_ms * 1e-3) / 1e9 / pipe_peak, "traffic": None, "peak_source": pipe_src,
                     "kernel_ms": k_ms, "required_alu_lane_ops": alu_ops, "required_xu_lane_ops": xu_ops,
                     "and_terms": n_terms,
                     "note": "REQUIRED work of the kernel's own formulation (per AND term and 8 words: 8 AND + 8 LOP3 "
                             "on the ALU pipe, 4 POPC on the XU pipe; row popcounts without the ANDs); the two pipes "
                             "run side by side, the slower one binds"}
    roofline = dict(roofline_hbm if binding[1] == "hbm" else roofline_pipe)
    roofline["min_time_us"] = {"hbm": t_hbm * 1e6, "alu": t_alu * 1e6, "xu": t_xu * 1e6}

    # ---- CPU baselines (child processes; bounded samples) ----
    cpu_baseline = None
    if not args.no_cpu_baseline:
        leg = cpu_leg("fitness", args, 12000, procs=1)
        if leg is not None:
            what = ("unmodified reference calculate_error from baseline/_ref" if leg["kind"] == "reference"
                    else "oracle port of calculate_error")
            cpu_baseline = {"value": leg["value"], "unit": UNIT, "cores": 1, "kind": leg["kind"],
                            "sample": f"1 individual x {args.nodes} nodes x {leg['cells']} cells = {leg['calls']} "
                                      f"calculate_error calls ({what}: Python eval per cell), {leg['seconds']:.1f} s, "
                                      f"single process like the reference"}
        if sweep is not None:
            leg = cpu_leg("sweep", args, 500)
            if leg is not None:
                what = ("unmodified reference CalculateImportanceScore.calculate_importance_scores() under its own "
                        "mp.Pool (baseline/_ref)" if leg["kind"] == "reference"
                        else "oracle C port, 64 cells per word, one thread per core")
                factor = args.cells * world / leg["cells"]
                sweep["cpu_baseline"] = {"value": leg["seconds"] * factor, "unit": "s (extrapolated)", "cores": leg["cores"],
                                         "kind": leg["kind"], "seconds_sample": leg["seconds"],
                                         "sample": f"{args.nodes} nodes x {leg['cells']} cells x 50 steps, "
                                                   f"{2 * args.nodes + 1} conditions ({what}); scale factor {factor:.0f} "
                                                   f"to {args.cells * world} cells"}
                mean_len = leg["mean_aligned_length"]
                # conds*steps*N*W LUT3 word-ops (full steps; the kernels stop a cell at its first
                # repeat, which the definition allows) + 2N*W*N*(mean aligned length) XOR+POPC word-ops
                lut_ops = (2 * args.nodes + 1) * 50.0 * args.nodes * n_words
                score_ops = 2.0 * args.nodes * n_words * args.nodes * mean_len
                sweep["roofline"] = {"kernel": "koki_plane_kernel + sweep_kernel<recheck, cleanup> (all passes)", "bound": "int_pipe_alu",
                                     "achieved": (lut_ops + score_ops) * world / sweep["seconds"] / 1e9,
                                     "peak": pipes["lop3"] * world, "unit": "Glane-op/s",
                                     "frac": (lut_ops + score_ops) / sweep["seconds"] / 1e9 / pipes["lop3"],
                                     "traffic": None, "peak_source": pipe_src,
                                     "algorithmic_lut3_word_ops": lut_ops, "algorithmic_scoring_word_ops": score_ops,
                                     "mean_aligned_length": mean_len,
                                     "note": "algorithmic ops = full 50 steps for every condition, one LUT3 per "
                                             "(node, 32 cells, step); the kernels use early exit (a CTA stops when its "
                                             "cells have repeated), cycle detection and deferral are not counted; what binds "
                                             "the dominant kernel (koki_plane_kernel, ncu, profiles/r2_sweep_final_ncu.txt): "
                                             "issue slots 77 %, ALU pipe 72 %, shared-memory wavefronts 75 % of peak"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args, {
                "exchange": ("none (1 GPU)" if world == 1 else
                             "peer memory: counts pushed over NVLink, summed inside the fitness kernel"
                             if exchange is not None else "ncclAllReduce(int64) between the kernels" + exchange_note),
                "timing": f"K steps back to back between one CUDA event pair; inputs larger than L2: step k reads "
                          f"copy k % {ga['n_rot']} of the bitplanes ({ga['n_rot']} x {ga['plane_bytes'] / 1e6:.1f} MB > 2 x 126 MB L2)",
                "wall_s_timed_region": ga["wall"], "fitness_checksum": ga["checksum"],
                "launch": "one cuda graph of K steps" if ga["graph"] else "eager",
                "overlap": "programmatic dependent launch both ways: the counting kernel of step k+1 claims its tiles "
                           "dynamically and runs on the SMs the fitness kernel of step k (three individuals per warp: "
                           "43 CTAs) leaves free; its flush waits for that kernel to complete",
                "ms_per_step_eager": ga["eager_ms"], "ms_per_step_isolated_l2_flushed": ga["isolated_ms"]}),
            "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_pipe": roofline_pipe,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "e2e_dense": e2e_dense, "sweep": sweep, "strong": strong,
            "clocks": clk, "gpu_launches": 2 * args.steps}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.cpu_leg:
        run_cpu_leg(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
