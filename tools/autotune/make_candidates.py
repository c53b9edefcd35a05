#!/usr/bin/env python
"""Autotuning of the pass lists of k_sim_paths by MEASURED time.

The SASS length of the loop bodies, which tools/pass_search minimises, predicts the kernel time only to within a few
percent: ptxas' instruction ORDER (runs of fp64 instructions) matters as much as the count, and nothing static predicted it.
This script therefore produces many candidates whose loop bodies are short (random walks through the space of
shared-contribution pass lists, every accepted step within --slack instructions of the best known length), builds the
stand-alone timing harness (harness.cu, ~5 s each) for every one of them under tools/variants/autotune/, and writes an index;
tools/autotune/run_candidates.py then times them all on the B200 in one gpurun call.

    python tools/autotune/make_candidates.py [--count 120] [--slack 50] [--jobs 8] [--extra "-DMCTS_SEL_FULL=7"]
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools", "pass_search"))
from loopstat import loop_bodies  # noqa: E402

CLASS = {1: "C1", 3: "C1", 2: "AM", 13: "AM", 4: "F1", 5: "F1", 14: "F1", 17: "F1", 6: "MG", 7: "MG", 8: "MG2", 9: "P1", 10: "P1",
         15: "P1", 11: "P2", 12: "P2", 16: "S16", 18: "S18", 19: "S19", 20: "S20"}
SEEDS = [
    [[9, 10, 6, 15, 7], [14, 4, 5, 8, 17], [3, 18, 1, 12, 11], [2, 16, 19, 20, 13]],     # 1853, shipped
    [[3, 18, 11, 1, 12], [14, 4, 5, 8, 17], [6, 7, 19, 16, 20], [13, 9, 15, 2, 10]],     # 1870
    [[4, 5, 14, 20, 17], [13, 8, 2, 19, 16], [9, 10, 7, 6, 15], [3, 12, 1, 11, 18]],     # 1878
    [[3, 18, 1, 12, 11], [4, 5, 14, 17, 16], [9, 10, 13, 15, 2], [20, 8, 6, 19, 7]],     # 1864
    [[4, 19, 5, 17, 14], [9, 10, 7, 15, 6], [13, 2, 8, 16, 20], [12, 3, 1, 11, 18]],     # 1883
]
ap = argparse.ArgumentParser()
ap.add_argument("--count", type=int, default=120)
ap.add_argument("--slack", type=int, default=50)
ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
ap.add_argument("--best", type=int, default=1853)
ap.add_argument("--extra", default="")
ap.add_argument("--out", default=os.path.join(ROOT, "tools", "variants", "autotune"))
args = ap.parse_args()
os.makedirs(args.out, exist_ok=True)
lock = threading.Lock()
kept, seen = {}, set()


def key_of(passes):
    return tuple(tuple(CLASS[i] for i in p) for p in passes)


def groups(p):
    g = {}
    for i in p:
        g.setdefault(CLASS[i], []).append(i)
    return g


def neighbour(passes, rng):
    for _ in range(200):
        q = [list(p) for p in passes]
        r = rng.random()
        if r < 0.45:
            p = rng.choice(q)
            a, b = rng.sample(range(5), 2)
            if CLASS[p[a]] == CLASS[p[b]]:
                continue
            p[a], p[b] = p[b], p[a]
        elif r < 0.65:
            a, b = rng.sample(range(4), 2)
            q[a], q[b] = q[b], q[a]
        else:
            a, b = rng.sample(range(4), 2)
            ga, gb = groups(q[a]), groups(q[b])
            ca = rng.sample(sorted(ga), rng.choice([1, 1, 2]) if len(ga) > 1 else 1)
            cb = rng.sample(sorted(gb), rng.choice([1, 1, 2]) if len(gb) > 1 else 1)
            ia, ib = [i for c in ca for i in ga[c]], [i for c in cb for i in gb[c]]
            if len(ia) != len(ib) or set(ca) == set(cb):
                continue
            pa, pb = [q[a].index(i) for i in ia], [q[b].index(i) for i in ib]
            for k in range(len(ia)):
                q[a][pa[k]], q[b][pb[k]] = ib[k], ia[k]
        if all(len(p) == 5 for p in q) and all(len({j for j, p in enumerate(q) if any(CLASS[i] == c for i in p)}) == 1 for c in set(CLASS.values())):
            return q
    return passes


def build(tag, passes):
    """Compile the harness for `passes`; returns the loop-body lengths (or None) and leaves the binary in args.out."""
    hdr, exe = os.path.join(args.out, tag + ".h"), os.path.join(args.out, tag)
    with open(hdr, "w") as f:
        f.write("".join("#define MCTS_FP%d %s\n" % (j, ", ".join(map(str, p))) for j, p in enumerate(passes)))
    cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-I", os.path.join(ROOT, "mc_trade_sizing_b200", "csrc"),
           "-I", os.path.join(ROOT, "include"), "-include", hdr] + args.extra.split() + ["-o", exe, os.path.join(HERE, "harness.cu")]
    if subprocess.run(cmd, capture_output=True).returncode:
        return None
    bodies = [n for n, _ in loop_bodies(exe)]
    return bodies if len(bodies) == 4 else None


def walker(wid):
    rng = random.Random(4242 + wid)
    cur = SEEDS[wid % len(SEEDS)]
    while True:
        with lock:
            if len(kept) >= args.count:
                return
        nb = neighbour(cur, rng) if rng.random() < 0.97 else SEEDS[rng.randrange(len(SEEDS))]
        k = key_of(nb)
        with lock:
            if k in seen:
                cur = nb if k in kept else cur
                continue
            seen.add(k)
            tag = "c%04d" % len(seen)
        bodies = build(tag, nb)
        ok = bodies is not None and sum(bodies) <= args.best + args.slack
        with lock:
            if ok and len(kept) < args.count:
                kept[k] = {"tag": tag, "passes": nb, "bodies": bodies, "sass": sum(bodies)}
                print("%s  %d %s  %s" % (tag, sum(bodies), bodies, nb), flush=True)
        if ok:
            cur = nb
        else:
            for ext in ("", ".h"):
                try:
                    os.remove(os.path.join(args.out, tag + ext))
                except OSError:
                    pass


with ThreadPoolExecutor(args.jobs) as ex:
    list(ex.map(walker, range(args.jobs)))
with open(os.path.join(args.out, "index.json"), "w") as f:
    json.dump(sorted(kept.values(), key=lambda c: c["sass"]), f, indent=1)
print("%d candidates kept of %d compiled -> %s" % (len(kept), len(seen), args.out))
