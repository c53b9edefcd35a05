#!/usr/bin/env python
"""Pass search for the shared-contribution layout of k_sim_paths (compile only, no GPU).

Ids of one behaviour class share a pass (their contribution select is computed once per trade, every id keeps its own
chains), so a pass is a multiset of classes and a composition is a partition of the 20 ids into four passes of five
that keeps every class whole.  For every pass that occurs in some partition this script compiles that pass ALONE
(one_pass.cu, ~2 s) in every distinct order of its classes, scores it by the SASS length of its 8-trade loop body
(loopstat.py), and then ranks the partitions by the sum of their passes' best scores.  The winners are confirmed in
the full kernel (one_full.cu) and on the GPU before they go into csrc/eval_core.cuh (MCTS_FP0..3).

    python tools/pass_search/search_shared.py [--sel 7] [--jobs 8] [--top 5]
"""
import argparse
import itertools
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from loopstat import loop_bodies  # noqa: E402

MEMBERS = {"C1": [1, 3], "AM": [2, 13], "F1": [4, 5, 14, 17], "MG": [6, 7], "MG2": [8], "P1": [9, 10, 15], "P2": [11, 12],
           "S16": [16], "S18": [18], "S19": [19], "S20": [20]}
ap = argparse.ArgumentParser()
ap.add_argument("--sel", type=int, default=7)
ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
ap.add_argument("--top", type=int, default=5)
ap.add_argument("--cache", default=os.path.join(tempfile.gettempdir(), "pass_search_shared.json"))
args = ap.parse_args()
tmp = tempfile.mkdtemp(prefix="pass_shared_")


def partitions():
    """All ways to put the 11 classes (kept whole) into four unordered bins of five ids."""
    names = sorted(MEMBERS, key=lambda c: -len(MEMBERS[c]))
    out = set()

    def rec(i, bins):
        if i == len(names):
            if all(sum(len(MEMBERS[c]) for c in b) == 5 for b in bins):
                out.add(tuple(sorted(tuple(sorted(b)) for b in bins)))
            return
        seen = set()
        for b in bins:
            key = tuple(sorted(b))
            if key in seen:
                continue
            seen.add(key)
            if sum(len(MEMBERS[c]) for c in b) + len(MEMBERS[names[i]]) <= 5:
                b.append(names[i])
                rec(i + 1, bins)
                b.pop()

    rec(0, [[], [], [], []])
    return sorted(out)


def orders(classes):
    """Distinct orders of a pass: permutations of the class-name multiset (ids of one class are interchangeable)."""
    slots = [c for c in classes for _ in MEMBERS[c]]
    return sorted(set(itertools.permutations(slots)))


def ids_of(order):
    it = {c: iter(MEMBERS[c]) for c in set(order)}
    return [next(it[c]) for c in order]


def compile_pass(job):
    tag, order = job
    hdr, cub = os.path.join(tmp, tag + ".h"), os.path.join(tmp, tag + ".cubin")
    with open(hdr, "w") as f:
        f.write("#define MCTS_ONE_PASS %s\n#define MCTS_SEL_FULL %d\n" % (", ".join(map(str, ids_of(order))), args.sel))
    cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-I", os.path.join(ROOT, "mc_trade_sizing_b200", "csrc"),
           "-include", hdr, "-cubin", "-o", cub, os.path.join(HERE, "one_pass.cu")]
    if subprocess.run(cmd, capture_output=True).returncode:
        return order, None
    bodies = loop_bodies(cub)
    os.remove(cub)
    return order, (bodies[0][0] if len(bodies) == 1 else None)


parts = partitions()
passes = sorted({p for part in parts for p in part})
cache = json.load(open(args.cache)) if os.path.exists(args.cache) else {}
jobs = []
for p in passes:
    for o in orders(p):
        key = "%d|%s" % (args.sel, ",".join(o))
        if key not in cache:
            jobs.append((key.replace("|", "_").replace(",", "_"), o))
print("%d partitions, %d distinct passes, %d orders to compile" % (len(parts), len(passes), len(jobs)), flush=True)
with ThreadPoolExecutor(args.jobs) as ex:
    for n, (order, score) in enumerate(ex.map(compile_pass, jobs)):
        cache["%d|%s" % (args.sel, ",".join(order))] = score
        if n % 50 == 49:
            json.dump(cache, open(args.cache, "w"))
            print("  %d / %d" % (n + 1, len(jobs)), flush=True)
json.dump(cache, open(args.cache, "w"))

best = {}
for p in passes:
    cands = [(cache.get("%d|%s" % (args.sel, ",".join(o))), o) for o in orders(p)]
    cands = [c for c in cands if c[0] is not None]
    best[p] = min(cands) if cands else (10 ** 9, None)
ranked = sorted(parts, key=lambda part: sum(best[p][0] for p in part))
for part in ranked[: args.top]:
    total = sum(best[p][0] for p in part)
    print("total %d instructions per 8 trades x 20 ids = %.2f per trade-step" % (total, total / 160.0))
    for p in part:
        print("   %4d  %s" % (best[p][0], ", ".join(map(str, ids_of(best[p][1])))))
