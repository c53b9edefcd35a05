#!/usr/bin/env python
"""BASELINE.json's configs through the UNCHANGED public API on every B200 of the box from ONE process
(MCTS_DEVICES=all -> mcts_group_*; no torchrun):   MCTS_DEVICES=all python tools/group_configs.py"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_trade_sizing_b200 import trading_models as tm  # noqa: E402

REG = [dict(length=300, hit_rate=0.9, avg_win=200, avg_loss=100), dict(length=200, hit_rate=0.5, avg_win=100, avg_loss=100),
       dict(length=500, hit_rate=0.2, avg_win=100, avg_loss=200)]
CONFIGS = [
    ("C3 Markov-2, 1e6 paths", (0.82, 186.0, 219.0, 1000, 50000, 20), dict(use_markov2=True, p_win_ww=0.85, p_win_wl=0.65, p_win_lw=0.55, p_win_ll=0.30)),
    ("C4 regime switching, 1e7 paths", (0.82, 186.0, 219.0, 1000, 500000, 20), dict(use_regime=True, regimes=REG)),
    ("fixed 67,109,120-path problem (strong scaling)", (0.82, 186.0, 219.0, 1000, 3355456, 20), {}),
    ("C5 scale sweep, 1e9 paths", (0.82, 186.0, 219.0, 1000, 50_000_000, 20), {}),
]
eng = tm.get_engine()
print("devices:", getattr(eng, "devices", [getattr(eng, "device", 0)]))
tm.run_all_strategies(0.82, 186.0, 219.0, 1000, 100000, 20, seed=1)      # contexts, buffers
for name, args, kw in CONFIGS:
    tm.run_all_strategies(*(args[:4] + (8 * max(len(getattr(eng, "devices", [0])), 1) * 1024, args[5])), seed=1, **kw)   # loads this mode's kernels on every device
    t0 = time.perf_counter()
    rows = tm.run_all_strategies(*args, seed=2026, **kw)
    wall = time.perf_counter() - t0
    by = {r[0]: r for r in rows}
    c = by["Constant position size 1"]
    sha = hashlib.sha256(np.array([by[tm.DESCRIPTIONS[i]][1:] for i in range(1, 21)], dtype=np.float64).tobytes()).hexdigest()
    print("%-48s wall %9.3f ms  %.3e trade-steps/s  avg_profit[1] = %.3f  avg_dd[1] = %.3f  summary sha256 %s"
          % (name, wall * 1e3, args[4] * args[5] * 1000 * 20 / wall, c[1], c[2], sha[:16]), flush=True)
