#!/usr/bin/env python
"""Smallest case for compute-sanitizer: one replay of 300 paths x 100 trades (three blocks, partial last block, four tiles with a
partial last tile, cp.async double buffering) and one 300-path simulate, checked against the oracle.
    compute-sanitizer --tool racecheck python tools/sanitize_replay.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_trade_sizing_b200 import Engine, make_params  # noqa: E402
from oracle import oracle  # noqa: E402

eng = Engine(0)
rng = np.random.default_rng(1)
vals = rng.choice([204.60000000000002, 167.4, -240.9, 186.0, -219.0], size=(300, 100))
out = eng.replay(vals)
ref = oracle.eval_paths(vals)
assert (out["profit"].view(np.uint64) == ref["profit"].view(np.uint64)).all() and (out["dd"].view(np.uint64) == ref["dd"].view(np.uint64)).all()
prm = make_params(0.82, 186.0, 219.0, 100, 30, 10)
sim = eng.simulate(prm, seed=3, per_path=True)
ref2 = oracle.eval_paths(eng.sequences(prm, seed=3), num_trades=100)
assert (sim["profit"].view(np.uint64) == ref2["profit"].view(np.uint64)).all()
print("sanitize_replay ok")
