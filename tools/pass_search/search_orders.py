#!/usr/bin/env python
"""Search over pass compositions of k_sim_paths by what ptxas makes of them (compile only, no GPU).

Which strategy ids share a pass, and in which order they are written, decides how ptxas allocates the register pairs
of the unrolled 8-trade loop body: the same five rules compile to 12.0 - 16 instructions per trade-step.  This script
scores a composition by the SASS length of the loop bodies (loopstat.py) and improves it by hill climbing (swap two
ids inside a pass / between passes / move an id; no two ids of one behaviour class may share a pass in full mode).
The winners are confirmed on the GPU (tools/build_variant.py + tools/profile_step.py) before they are written into
csrc/eval_core.cuh (MCTS_FP0..3, MCTS_DP0..2) and csrc/kernels.cuh (MCTS_SEL_*).

    python tools/pass_search/search_orders.py full  [--seconds 600] [--sel 7]
    python tools/pass_search/search_orders.py dedup [--seconds 600] [--sel 7]
"""
import argparse
import os
import random
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from loopstat import loop_bodies  # noqa: E402

CLASS = {1: "C1", 3: "C1", 2: "AM", 13: "AM", 4: "F1", 5: "F1", 14: "F1", 17: "F1", 6: "MG", 7: "MG", 8: "MG2", 9: "P1", 10: "P1",
         15: "P1", 11: "P2", 12: "P2", 16: "S16", 18: "S18", 19: "S19", 20: "S20"}
START = {"full": [[6, 9, 1, 2, 4], [7, 10, 3, 13, 5], [14, 11, 16, 15, 18], [17, 12, 8, 19, 20]],
         "dedup": [[18, 4, 1], [9, 8, 16, 6], [11, 20, 2, 19]]}
ap = argparse.ArgumentParser()
ap.add_argument("mode", choices=["full", "dedup"])
ap.add_argument("--seconds", type=int, default=600)
ap.add_argument("--sel", type=int, default=7, help="select mode of the kernel flavour (kernels.cuh MCTS_SEL_*)")
ap.add_argument("--chains", type=int, default=8)
args = ap.parse_args()
tmp = tempfile.mkdtemp(prefix="pass_search_")
cache = {}


def score(tag, passes):
    key = tuple(tuple(CLASS[i] for i in p) for p in passes)       # ids of one class compile to the same code
    if key in cache:
        return cache[key]
    hdr, cub = os.path.join(tmp, tag + ".h"), os.path.join(tmp, tag + ".cubin")
    prefix = "MCTS_FP" if args.mode == "full" else "MCTS_DP"
    with open(hdr, "w") as f:
        f.write("".join("#define %s%d %s\n" % (prefix, k, ", ".join(map(str, p))) for k, p in enumerate(passes)))
        f.write("#define MCTS_SEL_%s %d\n" % (args.mode.upper(), args.sel))
    cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-I", os.path.join(ROOT, "mc_trade_sizing_b200", "csrc"),
           "-include", hdr, "-cubin", "-o", cub, os.path.join(HERE, "one_%s.cu" % args.mode)]
    if subprocess.run(cmd, capture_output=True).returncode:
        cache[key] = None
        return None
    bodies = loop_bodies(cub)
    cache[key] = (sum(n for n, _ in bodies), bodies) if len(bodies) == len(passes) else None
    return cache[key]


def valid(passes):
    if args.mode == "dedup":
        return all(3 <= len(p) <= 5 for p in passes)
    return all(len(set(CLASS[i] for i in p)) == len(p) == 5 for p in passes)


def neighbour(passes, rng):
    for _ in range(100):
        q = [list(p) for p in passes]
        r = rng.random()
        if r < 0.45:
            p = rng.choice(q)
            a, b = rng.sample(range(len(p)), 2)
            p[a], p[b] = p[b], p[a]
        elif r < 0.9 or args.mode == "full":
            pa, pb = rng.sample(range(len(q)), 2)
            a, b = rng.randrange(len(q[pa])), rng.randrange(len(q[pb]))
            q[pa][a], q[pb][b] = q[pb][b], q[pa][a]
        else:
            pa, pb = rng.sample(range(len(q)), 2)
            q[pb].insert(rng.randrange(len(q[pb]) + 1), q[pa].pop(rng.randrange(len(q[pa]))))
        if valid(q):
            return q
    return passes


def chain(cid):
    rng = random.Random(cid)
    cur, t0 = START[args.mode], time.time()
    cur_s = score("c%d" % cid, cur)
    best = (cur_s, cur)
    while time.time() - t0 < args.seconds:
        nb = neighbour(cur, rng)
        s = score("c%d" % cid, nb)
        if s is not None and s[0] <= cur_s[0] + (4 if rng.random() < 0.1 else 0):
            cur, cur_s = nb, s
            if cur_s[0] < best[0][0]:
                best = (cur_s, cur)
                print("chain %d: %d instructions per 8 trades %s  %s" % (cid, cur_s[0], cur_s[1], cur), flush=True)
    return best


with ThreadPoolExecutor(args.chains) as ex:
    results = sorted(ex.map(chain, range(args.chains)), key=lambda r: r[0][0])
print("start:", score("start", START[args.mode]))
for s, p in results[:3]:
    print("best: %d %s %s" % (s[0], s[1], p))
