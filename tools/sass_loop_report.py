#!/usr/bin/env python
"""SASS evidence for the hot loops of libmcts.so (cuobjdump -sass on the shipped library, no GPU needed):

    python tools/sass_loop_report.py [path/to/libmcts.so] > profiles/r02_sass_loop_body.txt

For every kernel flavour of k_sim_paths / k_sim_paths_int / k_replay_paths it lists the innermost 8-trade loop bodies
(backward branches whose body holds 24 DADDs per strategy, or the VIADDMNMX pattern of the int32 evaluator) with
their instruction counts per pipe and per trade-step, and prints the full listing of the bodies of the headline
kernel k_sim_paths<16, PM_FULL, 128>.  Pipe attribution (B200): FP64 = D*, FMA-heavy = IMAD*, ALU = LOP3 / SHF / SEL /
FSEL / ISETP / IADD3 / VIADD* / VIMNMX / LEA / MOV / PRMT / PLOP3, LSU = LDS / STS / LDG / STG / LDL / STL."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mc_trade_sizing_b200", "libmcts.so")
PIPES = (("FP64", ("DADD", "DSETP", "DMUL", "DFMA")), ("FMA-heavy", ("IMAD",)),
         ("ALU", ("LOP3", "SHF", "SEL", "FSEL", "ISETP", "IADD3", "VIADD", "VIMNMX", "VIADDMNMX", "LEA", "MOV", "PRMT", "PLOP3", "IABS", "POPC")),
         ("LSU", ("LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMS", "LDSM")))


def pipe_of(op):
    for name, ops in PIPES:
        if op in ops:
            return name
    return "other"


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fn, funcs = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            fn = m.group(1)
            funcs[fn] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and fn:
            funcs[fn].append((int(m.group(1), 16), m.group(2).strip()))
    return funcs


def demangle(names):
    out = subprocess.run(["cu++filt"] + list(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def loops(ins):
    res = []
    for addr, text in ins:
        m = re.search(r"0x([0-9a-f]+)", text) if "BRA" in text else None
        if not m or int(m.group(1), 16) >= addr:
            continue
        lo = int(m.group(1), 16)
        seg = [(a, t) for a, t in ins if lo <= a <= addr]
        if any(re.search(r"0x([0-9a-f]+)", t) and "BRA" in t and lo <= int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) < a and a != addr for a, t in seg):
            continue  # not innermost
        n_dadd = sum(re.sub(r"^@!?U?P\d+\s+", "", t).startswith("DADD") for _, t in seg)   # three per fp64 trade-step
        n_mnmx = sum("VIADDMNMX" in t for _, t in seg)                                       # one per int32 trade-step
        if (n_dadd >= 72 and n_dadd % 24 == 0) or n_mnmx >= 80:
            res.append((seg, n_dadd // 24 if n_dadd else n_mnmx // 8))
    return res


def histogram(seg):
    ops, pipes, pred = collections.Counter(), collections.Counter(), 0
    for _, t in seg:
        if t.startswith("@"):
            pred += 1
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
        ops[op] += 1
        pipes[pipe_of(op)] += 1
    return ops, pipes, pred


funcs = functions(LIB)
names = demangle(list(funcs))
print("SASS loop bodies of %s (cuobjdump -sass); one body = 8 trades x NS strategies\n" % os.path.relpath(LIB, ROOT))
headline = []
for fn, ins in funcs.items():
    nice = names[fn].replace("mcts::", "").replace("(int)", "").replace("(bool)", "")
    if not any(k in nice for k in ("k_sim_paths", "k_replay_paths")):
        continue
    if re.search(r"<\d+, \d+, (32|64)[,>]", nice):
        continue  # 32- and 64-thread instantiations of the same code (used for very long paths)
    found = loops(ins)
    if not found:
        continue
    print("%s   (%d SASS instructions in all)" % (nice, len(ins)))
    total = 0
    for seg, ns in found:
        ops, pipes, pred = histogram(seg)
        steps = 8.0 * ns
        total += len(seg)
        print("  body @0x%05x: %3d instructions, %2d strategies -> %5.2f per trade-step | " % (seg[0][0], len(seg), ns, len(seg) / steps)
              + "  ".join("%s %.2f" % (p, pipes[p] / steps) for p in ("FP64", "FMA-heavy", "ALU", "LSU", "other")) + " | predicated %.2f" % (pred / steps))
        print("      " + ", ".join("%s %d" % kv for kv in ops.most_common()))
    n_ids = sum(ns for _, ns in found)
    print("  all bodies: %d instructions per 8 trades of %d strategy slots = %.2f per trade-step\n" % (total, n_ids, total / (8.0 * n_ids)))
    if "k_sim_paths<16, 0, 128, 0>" in nice:
        headline = found
if headline:
    print("=" * 110)
    print("Full listing of the loop bodies of k_sim_paths<16, 0, 128, 0> (the headline kernel: all 20 ids, fp64, Fisher-Yates generator)\n")
    for k, (seg, ns) in enumerate(headline):
        print("---- pass %d: %d instructions ----" % (k, len(seg)))
        for a, t in seg:
            print("  /*%05x*/  %s" % (a, t))
        print()
