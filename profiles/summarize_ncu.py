#!/usr/bin/env python
"""Turn an ncu report (gpurun_out/*.ncu-rep) into the small text summary committed here.

    python profiles/summarize_ncu.py gpurun_out/prof_r1b.ncu-rep > profiles/r1_ncu_full_summary.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = ncu_csv(rep, "--page", "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# ncu --set full summary of {rep} (per launch; cold-cache, serialised replays)")
    for d in data:
        print("\n## " + d[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print(f"{k:90s} {d[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        r = float(d[hdr.index("dram__bytes_read.sum")])
        w = float(d[hdr.index("dram__bytes_write.sum")])
        t = float(d[hdr.index("gpu__time_duration.sum")])
        ur, ut = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("gpu__time_duration.sum")]
        print(f"{'dram traffic read+write':90s} {r + w:16.1f} {ur}  over {t:.1f} {ut}")
    src = ncu_csv(rep, "--page", "source")
    cur, ie = None, None
    hist = {}
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = r[1]
            hist.setdefault(cur, Counter())
            continue
        if r and r[0] == "Address":
            ie = r.index("Instructions Executed")
            continue
        if cur and ie is not None and len(r) > ie:
            toks = r[1].strip().split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            try:
                hist[cur][op.split(".")[0]] += int(r[ie])
            except ValueError:
                pass
    for k, c in hist.items():
        tot = sum(c.values())
        if not tot:
            continue
        print(f"\n## SASS opcode mix (executed warp instructions, all profiled launches): {k}")
        print("  " + ", ".join(f"{op} {n / tot * 100:.1f}%" for op, n in c.most_common(16)))


if __name__ == "__main__":
    main(sys.argv[1])
