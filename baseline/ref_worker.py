#!/usr/bin/env python
"""One worker of the Python-reference CPU baseline: imports the UNMODIFIED reference module from baseline/_ref (by
path, read-only), seeds NumPy's global RandomState the way BASELINE.md section 3 prescribes, runs its
run_all_strategies in process on this core and prints one JSON line with the time spent inside the call.

    python baseline/ref_worker.py MODE SIMS SHUFFLES SEED [START_AT_EPOCH_SECONDS]

MODE: shuffle | markov1 | markov2 | regime  (parameters of BASELINE.json configs C1-C4, 1000 trades)."""
import importlib.util
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = {
    "shuffle": {},
    "markov1": dict(use_markov=True, p_win_after_win=0.75, p_win_after_loss=0.60),
    "markov2": dict(use_markov2=True, p_win_ww=0.85, p_win_wl=0.65, p_win_lw=0.55, p_win_ll=0.30),
    "regime": dict(use_regime=True, regimes=[dict(length=300, hit_rate=0.9, avg_win=200, avg_loss=100),
                                             dict(length=200, hit_rate=0.5, avg_win=100, avg_loss=100),
                                             dict(length=500, hit_rate=0.2, avg_win=100, avg_loss=200)]),
}


def main():
    mode, sims, shuffles, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    start_at = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
    import numpy as np
    spec = importlib.util.spec_from_file_location("ref_trading_models", os.path.join(HERE, "_ref", "src", "trading_models.py"))
    tm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tm)
    while time.time() < start_at:      # all workers of a group start together, after their imports
        time.sleep(0.001)
    np.random.seed(seed)
    t0 = time.perf_counter()
    rows = tm.run_all_strategies(0.82, 186.0, 219.0, 1000, sims, shuffles, **MODES[mode])
    secs = time.perf_counter() - t0
    const = next(r for r in rows if r[0].startswith("Constant position size 1"))
    print(json.dumps({"mode": mode, "seconds": secs, "trade_steps": sims * shuffles * 1000 * 20, "avg_profit_strategy1": float(const[1])}))


if __name__ == "__main__":
    main()
