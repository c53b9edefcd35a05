#!/usr/bin/env python
"""In-situ refinement of the shared-contribution pass lists of k_sim_paths (compile only, no GPU).

search_shared.py scores every candidate pass compiled ALONE; inside the four-pass kernel ptxas allocates the same loop
body differently (the position of a pass in the kernel changes its register-pair moves by up to 15 %).  This script
therefore hill-climbs on the real thing: one_full.cu (the whole hot kernel) is compiled for every candidate and scored
by the summed SASS length of its four 8-trade loop bodies.  Moves keep every behaviour class whole inside one pass:
swap two ids inside a pass, swap two passes, exchange equal-sized class groups between passes.

    python tools/pass_search/search_insitu.py [--seconds 900] [--chains 7] [--sel 7] [--noinline 0|1] [--start "a;b;c;d"]
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
STARTS = [
    [[9, 10, 15, 6, 7], [1, 3, 11, 12, 18], [8, 4, 5, 14, 17], [2, 13, 19, 20, 16]],
    [[9, 10, 15, 6, 7], [1, 3, 11, 12, 18], [4, 5, 14, 17, 19], [2, 13, 8, 20, 16]],
    [[9, 10, 15, 2, 13], [1, 3, 11, 12, 18], [4, 5, 14, 17, 19], [6, 7, 20, 8, 16]],
    [[9, 10, 15, 6, 7], [1, 3, 11, 12, 18], [4, 5, 14, 17, 16], [2, 13, 20, 8, 19]],
    [[9, 10, 15, 2, 13], [1, 3, 11, 12, 18], [8, 4, 5, 14, 17], [6, 20, 19, 16, 7]],
]
ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=int, default=900)
ap.add_argument("--chains", type=int, default=7)
ap.add_argument("--sel", type=int, default=7)
ap.add_argument("--noinline", type=int, default=0)
ap.add_argument("--start", default="")
ap.add_argument("--extra", default="", help="extra #define lines for the candidate header, e.g. 'MCTS_NEST 1;MCTS_STEP_ORDER 1'")
ap.add_argument("--kernel", default="full", choices=["full", "replay"], help="full: k_sim_paths (MCTS_FP*, MCTS_SEL_FULL); replay: k_replay_paths (MCTS_RP*, MCTS_SEL_REPLAY)")
args = ap.parse_args()
if args.start:
    STARTS = [[[int(x) for x in p.split(",")] for p in args.start.split(";")]]
tmp = tempfile.mkdtemp(prefix="pass_insitu_")
cache = {}


def key_of(passes):
    return tuple(tuple(CLASS[i] for i in p) for p in passes)   # ids of one class compile to the same code


def score(tag, passes):
    k = key_of(passes)
    if k in cache:
        return cache[k]
    hdr, cub = os.path.join(tmp, tag + ".h"), os.path.join(tmp, tag + ".cubin")
    with open(hdr, "w") as f:
        prefix, selname = ("MCTS_FP", "MCTS_SEL_FULL") if args.kernel == "full" else ("MCTS_RP", "MCTS_SEL_REPLAY")
        f.write("".join("#define %s%d %s\n" % (prefix, j, ", ".join(map(str, p))) for j, p in enumerate(passes)))
        f.write("#define %s %d\n#define MCTS_NOINLINE_PASSES %d\n" % (selname, args.sel, args.noinline))
        f.write("".join("#define %s\n" % d.strip() for d in args.extra.split(";") if d.strip()))
    cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-I", os.path.join(ROOT, "mc_trade_sizing_b200", "csrc"),
           "-include", hdr, "-cubin", "-o", cub, os.path.join(HERE, "one_%s.cu" % args.kernel)]
    if subprocess.run(cmd, capture_output=True).returncode:
        cache[k] = None
        return None
    bodies = loop_bodies(cub)
    cache[k] = (sum(n for n, _ in bodies), [n for n, _ in bodies]) if len(bodies) == 4 else None
    return cache[k]


def groups(p):
    g = {}
    for i in p:
        g.setdefault(CLASS[i], []).append(i)
    return g


def neighbour(passes, rng):
    for _ in range(200):
        q = [list(p) for p in passes]
        r = rng.random()
        if r < 0.5:      # order inside a pass
            p = rng.choice(q)
            a, b = rng.sample(range(5), 2)
            if CLASS[p[a]] == CLASS[p[b]]:
                continue
            p[a], p[b] = p[b], p[a]
        elif r < 0.65:   # position of two passes in the kernel
            a, b = rng.sample(range(4), 2)
            q[a], q[b] = q[b], q[a]
        else:            # exchange equal-sized sets of whole classes between two passes
            a, b = rng.sample(range(4), 2)
            ga, gb = groups(q[a]), groups(q[b])
            ca = rng.sample(sorted(ga), rng.choice([1, 1, 2]) if len(ga) > 1 else 1)
            cb = rng.sample(sorted(gb), rng.choice([1, 1, 2]) if len(gb) > 1 else 1)
            ia, ib = [i for c in ca for i in ga[c]], [i for c in cb for i in gb[c]]
            if len(ia) != len(ib) or set(ca) == set(cb):
                continue
            pos_a, pos_b = [q[a].index(i) for i in ia], [q[b].index(i) for i in ib]
            for k2 in range(len(ia)):
                q[a][pos_a[k2]], q[b][pos_b[k2]] = ib[k2], ia[k2]
        if all(len(p) == 5 for p in q) and all(len({j for j, p in enumerate(q) if any(CLASS[i] == c for i in p)}) == 1 for c in set(CLASS.values())):
            return q
    return passes


def chain(cid):
    rng = random.Random(1000 + cid)
    cur, t0 = STARTS[cid % len(STARTS)], time.time()
    cur_s = score("c%d" % cid, cur)
    best = (cur_s, cur)
    print("chain %d start: %s %s" % (cid, cur_s, cur), flush=True)
    while time.time() - t0 < args.seconds:
        nb = neighbour(cur, rng)
        s = score("c%d" % cid, nb)
        if s is None:
            continue
        if cur_s is None or s[0] <= cur_s[0] + (6 if rng.random() < 0.15 else 0):
            cur, cur_s = nb, s
            if best[0] is None or cur_s[0] < best[0][0]:
                best = (cur_s, cur)
                print("chain %d: %d %s  %s" % (cid, cur_s[0], cur_s[1], cur), flush=True)
    return best


with ThreadPoolExecutor(args.chains) as ex:
    results = sorted((r for r in ex.map(chain, range(args.chains)) if r[0] is not None), key=lambda r: r[0][0])
for s, p in results[:5]:
    print("best: %d = %.2f per trade-step %s  %s" % (s[0], s[0] / 160.0, s[1], "; ".join(",".join(map(str, x)) for x in p)))
