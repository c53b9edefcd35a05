"""The CLI shim must print byte-for-byte what the reference's main() prints for the same result rows
(tests/golden/cli.json was captured from the live reference with run_all_strategies stubbed), and the
unchanged dps.py scraper (regexes restated from dps.py:145-154 and dps.py:356) must parse it."""
import contextlib
import io
import json
import os
import re
import sys

import numpy as np

from conftest import GOLDEN
from mc_trade_sizing_b200 import trading_models as tm

ROW_RE = (r"(.+?)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)"
          r"\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)\s+(-?\d+\.\d+)")   # dps.py:356


def _run_main(argv, table, monkeypatch):
    monkeypatch.setattr(tm, "run_all_strategies", lambda *a, **k: table)
    monkeypatch.setitem(sys.modules, "colorama", None)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tm.main(argv)
    return buf.getvalue()


def test_stdout_is_byte_identical_to_the_reference(kat, monkeypatch):
    golden = json.load(open(os.path.join(GOLDEN, "cli.json")))
    for name, case in golden.items():
        ids, rows = kat[name + "/rank_ids"], kat[name + "/rows"]
        table = [(tm.DESCRIPTIONS[int(i)],) + tuple(np.float64(x) for x in r) for i, r in zip(ids, rows)]
        assert _run_main(case["argv"], table, monkeypatch) == case["stdout"], name


def test_reference_scraper_parses_our_output(kat, monkeypatch):
    golden = json.load(open(os.path.join(GOLDEN, "cli.json")))
    case = golden["a5_markov1"]
    ids, rows = kat["a5_markov1/rank_ids"], kat["a5_markov1/rows"]
    table = [(tm.DESCRIPTIONS[int(i)],) + tuple(np.float64(x) for x in r) for i, r in zip(ids, rows)]
    text = _run_main(case["argv"], table, monkeypatch)
    assert re.search(r"Hit rate: ([\d.]+)%", text).group(1) == "77.00"              # dps.py:147
    assert re.search(r"Mode: (.*?)\n", text).group(1) == "1st Order Markov"          # dps.py:148
    assert float(re.search(r"Average win per trade: €([\d.]+)", text).group(1)) == 186.0
    assert int(re.search(r"Number of shuffles per simulation: (\d+)", text).group(1)) == 2
    assert re.search(r"Break-even hit rate: ([\d.]+)%", text).group(1) == "56.24"    # :575-584 with the Markov-1 factor
    lines = [ln for ln in text.split("\n") if len(ln.split()) >= 10 and "Top 4 strategies compared to" not in ln]  # dps.py:345-348
    parsed = [re.match(ROW_RE, ln) for ln in lines]
    parsed = [m for m in parsed if m]
    assert len(parsed) == 20
    for m, row in zip(parsed, table):
        assert m.group(1).strip() == row[0]
        assert [float(m.group(i)) for i in range(2, 11)] == [float("%.2f" % x) for x in row[1:]]


def test_flags_and_defaults_match_the_reference_parser():
    a = tm.build_parser().parse_args(["--hit_rate", "0.7", "--avg_win", "1", "--avg_loss", "2"])
    assert (a.num_simulations, a.num_trades, a.num_mc_shuffles) == (200, 400, 200)      # :537-539
    assert (a.p_win_after_win, a.p_win_after_loss) == (0.7, 0.5)
    assert (a.p_win_ww, a.p_win_wl, a.p_win_lw, a.p_win_ll) == (0.8, 0.6, 0.5, 0.3)
    assert not (a.use_markov or a.use_markov2 or a.use_regime) and a.regimes is None


def test_break_even_helper():
    assert tm.find_break_even_hit_rate(186.0, 219.0, "No Markov") == 219.0 / 405.0
    assert tm.find_break_even_hit_rate(186.0, 219.0, "2nd Order Markov") == 219.0 / 405.0 * (1 + (0.8 - 0.5) / 4)
    assert tm.find_break_even_hit_rate(0, 219.0, "No Markov") is None


def test_sweep_builds_the_same_twelve_command_lines_as_dps(monkeypatch):
    """sweep.sweep_argvs mirrors dps.py:187-191, 211-259 (3 hit rates x 4 modes, str() formatting)."""
    from mc_trade_sizing_b200 import sweep
    cfg = {"hit_rate": 0.82, "avg_win": 186, "avg_loss": 219, "num_simulations": 200, "num_trades": 1000, "num_mc_shuffles": 20,
           "p_win_after_win": 0.8, "p_win_after_loss": 0.4, "p_win_ww": 0.8, "p_win_wl": 0.6, "p_win_lw": 0.5, "p_win_ll": 0.3,
           "regimes": [{"length": 300, "hit_rate": 0.9, "avg_win": 200, "avg_loss": 100}]}
    runs = sweep.sweep_argvs(cfg)
    assert len(runs) == 12
    assert [r[2][1] for r in runs[::4]] == ["0.72", "0.7699999999999999", "0.82"]          # SURVEY 3.1 [probe]
    assert runs[1][2][-5:] == ["--use_markov", "--p_win_after_win", "0.8", "--p_win_after_loss", "0.4"]
    assert runs[3][2][-3:-1] == ["--use_regime", "--regimes"] and json.loads(runs[3][2][-1]) == cfg["regimes"]
    parser = tm.build_parser()
    for _, _, argv in runs:
        parser.parse_args(argv)
    calls = []
    monkeypatch.setattr(tm, "run_all_strategies", lambda *a, **k: calls.append((a, k)) or [("x",) + (1.0,) * 9] * 20)
    out = sweep.run_dps_sweep(cfg)
    assert len(out) == 12 and len(calls) == 12 and all("Break-even hit rate" in o[3] for o in out)
    assert calls[3][1]["use_regime"] and calls[3][1]["regimes"] == cfg["regimes"] and calls[2][1]["use_markov2"]
