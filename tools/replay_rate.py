#!/usr/bin/env python
"""Kernel-side rate of the replay mode (k_replay_paths) on the library named by MCTS_LIB (default: the product's):
    python tools/replay_rate.py [paths] [trades]
prints trade-steps/s of the evaluation kernels of one mcts_replay call over `paths` caller-supplied sequences (C5 values)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from mc_trade_sizing_b200.engine import Engine, make_params  # noqa: E402

paths = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
trades = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = Engine(0)
seqs = np.ascontiguousarray(eng.sequences(make_params(0.82, 186.0, 219.0, trades, paths // 20, 20), seed=5))
best = None
for dedup in (False, True):
    rates = []
    for _ in range(4):
        out = eng.replay(seqs, per_path=False, options=eng.options(dedup=dedup))
        t = eng.timing()
        rates.append(t["trade_steps"] / (t["paths_ms"] * 1e-3))
    print("%s dedup=%d paths=%d trades=%d  kernels %s  (best %.4g trade-steps/s)  sha-ish %.17g" % (
        os.environ.get("MCTS_LIB", "product"), dedup, paths, trades, " ".join("%.4g" % r for r in rates), max(rates), float(out["summary"].sum())))
