#!/usr/bin/env python
"""Condense an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into per-kernel totals and shares.
    python tools/launch_summary.py gpurun_out/r01_bench_launches.csv"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = f.readlines()
start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
rd = csv.DictReader(lines[start:])
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("mcts::", "")
    tot[name][0] += 1
    tot[name][1] += ms
total = sum(v[1] for v in tot.values())
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("  %-40s launches %4d  total %10.3f ms  share %5.1f%%  avg %8.3f ms" % (name, n, ms, 100 * ms / total, ms / n))
