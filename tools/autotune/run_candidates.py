#!/usr/bin/env python
"""Time every candidate of tools/variants/autotune/index.json on the GPU (run under gpurun) and write the ranking.
    python tools/autotune/run_candidates.py [--rounds 2] > gpurun_out/autotune.txt"""
import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--dir", default=os.path.join(ROOT, "tools", "variants", "autotune"))
ap.add_argument("--rounds", type=int, default=2)
a = ap.parse_args()
index = json.load(open(os.path.join(a.dir, "index.json")))
best = {}
for rnd in range(a.rounds):      # interleaved repetitions: drift of the box averages out
    for c in index:
        exe = os.path.join(a.dir, c["tag"])
        try:
            out = subprocess.run([exe, c["tag"]], capture_output=True, text=True, timeout=120).stdout
        except Exception as e:
            out = "error %r" % (e,)
        m = re.search(r"([0-9.]+) ms", out)
        if m:
            best[c["tag"]] = min(best.get(c["tag"], 1e9), float(m.group(1)))
        else:
            print("# %s: %s" % (c["tag"], out.strip()[:200]), file=sys.stderr)
rows = sorted(((best.get(c["tag"], 1e9), c) for c in index), key=lambda r: r[0])
for ms, c in rows:
    print("%.4f ms  sass %d %s  %s" % (ms, c["sass"], c["bodies"], "; ".join(",".join(map(str, p)) for p in c["passes"])))
