#!/usr/bin/env python
"""Condense an .ncu-rep (first kernel) into the handful of metrics DESIGN.md / profiles/ quote.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def summarize(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    print("==", path, "--", d.get("Kernel Name", ("", ""))[1], "grid", d.get("Grid Size", ("", ""))[1], "block", d.get("Block Size", ("", ""))[1])
    for k in KEYS:
        if k in d:
            print("  %-78s %-10s %s" % (k, d[k][0], d[k][1]))
    stalls = sorted(((float(v[1]), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                     if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")), reverse=True)
    print("  stall reasons (warps per issue-active cycle):", ", ".join("%s %.2f" % (n, x) for x, n in stalls[:8]))


for p in sys.argv[1:]:
    summarize(p)
