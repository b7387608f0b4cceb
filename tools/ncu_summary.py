"""Summarise an .ncu-rep (``--set full --import-source on``) as text for profiles/: per kernel the pipe /
issue / shared-memory figures, the stall mix, the opcode histogram and the shared-memory wavefronts per opcode.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-name-substring] > profiles/rN_x_ncu.txt
"""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
       "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct"]


def ncu(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = ncu(rep, "--page", "raw", "--csv")
    hdr = rows[0]
    name_i = hdr.index("Kernel Name")
    print(f"# {rep}: ncu --set full --clock-control none --import-source on")
    for r in rows[2:]:
        if want and want not in r[name_i]:
            continue
        print(f"\n== {r[name_i]}")
        for k in RAW:
            if k in hdr:
                print(f"{k:86s} {r[hdr.index(k)]}")
        stalls = [(h, float(r[i])) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]]
        print("stall reasons (warps per issue-active cycle): " +
              ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}"
                        for h, v in sorted(stalls, key=lambda x: -x[1])[:8]))
    src = ncu(rep, "--page", "source", "--csv", "--print-source", "sass")
    kern = None
    data = collections.OrderedDict()
    for r in src:
        if not r:
            continue
        if r[0] == "Kernel Name":
            kern = r[1]
            data[kern] = {"hdr": None, "rows": []}
        elif r[0] == "Address" and kern:
            data[kern]["hdr"] = r
        elif kern and r[0].startswith("0x"):
            data[kern]["rows"].append(r)
    for kern, d in data.items():
        if (want and want not in kern) or not d["hdr"]:
            continue
        ix = {n: i for i, n in enumerate(d["hdr"])}
        ops = collections.Counter()
        wf = collections.defaultdict(lambda: [0, 0])
        for r in d["rows"]:
            t = r[ix["Source"]].split()
            op = t[1] if t[0].startswith("@") else t[0]
            n = int(r[ix["Instructions Executed"]])
            ops[op.split(".")[0]] += n
            w = int(r[ix["L1 Wavefronts Shared"]])
            if w:
                wf[op][0] += n
                wf[op][1] += w
        total = sum(ops.values())
        print(f"\n== opcode histogram (warp instructions), {kern}: {total}")
        print(", ".join(f"{o} {100 * v / total:.1f}%" for o, v in ops.most_common(14)))
        print("shared-memory wavefronts per opcode: " +
              ", ".join(f"{o} {w[1]} ({w[1] / max(w[0], 1):.2f}/inst)" for o, w in sorted(wf.items(), key=lambda x: -x[1][1])))


if __name__ == "__main__":
    main()
