#!/usr/bin/env python
"""Curve-export launch used for the HBM-bound ncu capture: 131,072 paths x 1000 trades x 20 strategies
= 21 GB of equity curves written to device memory."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_trade_sizing_b200 import Engine, make_params  # noqa: E402

eng = Engine(0)
prm = make_params(0.82, 186.0, 219.0, 1000, 200000, 20)
n = 1 << 17
for r in range(2):
    eng.export_curves(prm, 3 + r, 0, n, to_host=False)
    t = eng.timing()
    print("curve export %d paths: %.3f ms, %.1f GB/s, %.4g trade-steps/s" % (n, t["paths_ms"], t["trade_steps"] * 8 / t["paths_ms"] / 1e6,
                                                                               t["trade_steps"] / t["paths_ms"] * 1e3))
