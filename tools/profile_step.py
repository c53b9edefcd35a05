#!/usr/bin/env python
"""One chunk of the bench workload (C5 parameters, 2,097,140 paths) -- the command ncu wraps.
    python tools/profile_step.py [--dedup] [--mode dynamic|markov1|markov2|regime] [--sims N] [--reps R]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_trade_sizing_b200 import Engine, make_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dedup", action="store_true")
ap.add_argument("--mode", default="dynamic")
ap.add_argument("--sims", type=int, default=104857)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--trades", type=int, default=1000)
ap.add_argument("--shuffles", type=int, default=20)
ap.add_argument("--block", type=int, default=0)
ap.add_argument("--no-int", action="store_true")
a = ap.parse_args()
kw = {"dynamic": {}, "markov1": dict(use_markov=True, p_win_after_win=0.75, p_win_after_loss=0.60),
      "markov2": dict(use_markov2=True, p_win_ww=0.85, p_win_wl=0.65, p_win_lw=0.55, p_win_ll=0.30),
      "regime": dict(use_regime=True, regimes=[dict(length=300, hit_rate=0.9, avg_win=200, avg_loss=100),
                                               dict(length=200, hit_rate=0.5, avg_win=100, avg_loss=100),
                                               dict(length=500, hit_rate=0.2, avg_win=100, avg_loss=200)])}[a.mode]
eng = Engine(0)
prm = make_params(0.82, 186.0, 219.0, a.trades, a.sims, a.shuffles, **kw)
for r in range(a.reps):
    out = eng.simulate(prm, seed=100 + r, options=eng.options(chunk_paths=1 << 21, dedup=a.dedup, block_threads=a.block, int_fast_path=not a.no_int))
    t = eng.timing()
    print("rep %d: %.3f ms total, paths kernel %.3f ms, gen %.3f ms, reduce %.3f ms, %.4g trade-steps/s (kernel %.4g), avg_profit[1]=%.3f"
          % (r, t["total_ms"], t["paths_ms"], t["gen_ms"], t["reduce_ms"], t["trade_steps"] / t["total_ms"] * 1e3,
             t["trade_steps"] / t["paths_ms"] * 1e3, out["summary"][0][0]))

if os.environ.get("MCTS_PROFILE_REPLAY"):
    import time
    import numpy as np
    n = int(os.environ["MCTS_PROFILE_REPLAY"])
    seqs = eng.sequences(make_params(0.82, 186.0, 219.0, a.trades, n // 20, 20, **kw), seed=1)
    for r in range(2):
        t0 = time.perf_counter()
        eng.replay(seqs, per_path=False)
        wall = time.perf_counter() - t0
        t = eng.timing()
        print("replay %d paths: wall %.1f ms (%.3g trade-steps/s end to end, %.1f MB H2D), kernels %.3f ms (%.3g trade-steps/s)"
              % (seqs.shape[0], wall * 1e3, t["trade_steps"] / wall, t["h2d_bytes"] / 1e6, t["paths_ms"], t["trade_steps"] / t["paths_ms"] * 1e3))
