"""Orchestrator-level batching: the 12 runs `dps.py` launches as 12 subprocesses
(3 hit rates x {no Markov, Markov-1, Markov-2, regime}; reference dps.py:187-191, 211-259) executed in
ONE process on one engine -- one CUDA context instead of twelve.  Each run's text is exactly what the
CLI prints (`trading_models.main`), so `dps.py`'s scraper (dps.py:145-165, 339-383) can parse it.

    python -m mc_trade_sizing_b200.sweep path/to/dps_config.yaml          # prints the 12 blocks
    runs = run_dps_sweep(cfg_dict)                                         # [(argv, stdout text), ...]

`dps.py` itself stays unchanged; a deployment that wants the batching calls this module instead of
spawning the CLI twelve times.
"""
from __future__ import annotations

import contextlib
import io
import json
import sys

from . import trading_models as tm


def sweep_argvs(cfg: dict):
    """The 12 argument lists, built the way dps.py:214-259 builds them (same str() formatting, so
    e.g. the middle hit rate is '0.7699999999999999')."""
    base_hit = cfg["hit_rate"]
    out = []
    for hit_rate in (base_hit - 0.1, base_hit - 0.05, base_hit):
        cmd = ["--hit_rate", str(hit_rate), "--avg_win", str(cfg["avg_win"]), "--avg_loss", str(cfg["avg_loss"]),
               "--num_simulations", str(cfg["num_simulations"]), "--num_trades", str(cfg["num_trades"]),
               "--num_mc_shuffles", str(cfg["num_mc_shuffles"])]
        out.append(("without Markov", hit_rate, list(cmd)))
        out.append(("with Markov 1.Ord", hit_rate, cmd + ["--use_markov", "--p_win_after_win", str(cfg.get("p_win_after_win", 0.7)),
                                                          "--p_win_after_loss", str(cfg.get("p_win_after_loss", 0.5))]))
        out.append(("with Markov 2.Ord", hit_rate, cmd + ["--use_markov2", "--p_win_ww", str(cfg.get("p_win_ww", 0.8)),
                                                          "--p_win_wl", str(cfg.get("p_win_wl", 0.6)), "--p_win_lw", str(cfg.get("p_win_lw", 0.5)),
                                                          "--p_win_ll", str(cfg.get("p_win_ll", 0.3))]))
        reg = cmd + ["--use_regime"]
        if cfg.get("regimes"):
            reg += ["--regimes", json.dumps(cfg["regimes"])]
        out.append(("with Regime-Switching-Modell", hit_rate, reg))
    return out


def run_dps_sweep(cfg: dict):
    """Run the 12 simulations in this process; returns [(label, hit_rate, argv, stdout_text)]."""
    results = []
    for label, hit_rate, argv in sweep_argvs(cfg):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            tm.main(argv)
        results.append((label, hit_rate, argv, buf.getvalue()))
    return results


if __name__ == "__main__":
    import yaml
    with open(sys.argv[1]) as f:
        config = yaml.safe_load(f)
    for i, (label, hit_rate, argv, text) in enumerate(run_dps_sweep(config), start=1):
        print("#### run %d/12: hit rate %s %s" % (i, hit_rate, label))
        print(text)
