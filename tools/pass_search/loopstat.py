#!/usr/bin/env python
"""Instruction count of the 8-trade loop bodies of k_sim_paths in a cubin (cuobjdump -sass): for every innermost
backward branch whose body holds 24 x NS DADDs (three per trade-step) and no local-memory traffic, the body length and NS.
    python tools/pass_search/loopstat.py file.cubin"""
import re
import subprocess
import sys


def loop_bodies(cubin):
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    bodies = []
    for addr, text in ins:
        m = re.search(r"0x([0-9a-f]+)", text) if "BRA" in text else None
        if not m or int(m.group(1), 16) >= addr:
            continue
        seg = [t for a, t in ins if int(m.group(1), 16) <= a <= addr]
        n_dadd = sum(re.sub(r"^@!?U?P\d+\s+", "", t).startswith("DADD") for t in seg)   # 3 per trade-step
        if n_dadd >= 72 and n_dadd % 24 == 0 and not any(t.startswith(("LDL", "STL")) for t in seg):
            bodies.append((len(seg), n_dadd // 24))   # instructions per 8 trades, strategies in the pass
    return bodies


if __name__ == "__main__":
    b = loop_bodies(sys.argv[1])
    print(" | ".join("%d strategies: %d instr = %.2f per trade-step" % (ns, n, n / (8.0 * ns)) for n, ns in b),
          "| total per trade: %.1f" % (sum(n for n, _ in b) / 8.0))
