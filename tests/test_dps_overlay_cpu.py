"""The reference's UNCHANGED dps.py, run end to end against our CLI: 12 subprocesses of
src/trading_models.py, stdout scraped by dps.py's own regexes (dps.py:145-165, :356), deduped (:376-383) and written by
the reference's own sinks (src/output_handler.py).  Without a GPU the hot function is answered by the test oracle (the
CLI, the table format and everything downstream are the product's and the reference's); the GPU twin of this test
(test_dps_overlay_gpu.py) runs the real thing."""
import os

import pytest

import dps_overlay
from mc_trade_sizing_b200 import trading_models as tm

REF = dps_overlay.reference_dir()
pytestmark = pytest.mark.skipif(REF is None, reason="needs the reference checkout (/root/reference or baseline/_ref)")


def _small(path):
    """Shrink the Monte Carlo size only (50 x 4 instead of 200 x 20 paths): the CPU oracle answers twelve runs in seconds."""
    text = open(path, encoding="utf-8").read()
    text = text.replace("num_simulations: 200", "num_simulations: 50").replace("num_mc_shuffles: 20", "num_mc_shuffles: 4")
    open(path, "w", encoding="utf-8").write(text)


def test_unchanged_dps_scrapes_and_stores_our_table(tmp_path):
    ov = dps_overlay.build_overlay(str(tmp_path / "overlay"), REF, entry=dps_overlay.ENTRY_ORACLE, config_edit=_small)
    assert os.path.realpath(os.path.join(ov, "dps.py")) == os.path.join(REF, "dps.py")     # unchanged, not a copy
    p = dps_overlay.run_dps(ov, extra_env={"MCTS_SEED": "2026"})
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "Starting 12 simulations" in p.stdout and "All simulations completed." in p.stdout
    rows = dps_overlay.check_outputs(ov)
    # settings scraped from our header (dps.py:145-165): three hit rates x four modes, the strategy names of :451-472
    assert {r["Mode"] for r in rows} == {"No Markov", "1st Order Markov", "2nd Order Markov", "Regime Switching"}
    assert {r["Hit Rate (%)"] for r in rows} == {"72.00", "77.00", "82.00"}
    assert {r["Strategy"] for r in rows} == set(tm.DESCRIPTIONS.values())
    assert all(r["Num Simulations"] == 50 and r["Num Trades"] == 1000 and r["Num Shuffles"] == 4 for r in rows)
    # the equivalence classes of SURVEY A.2 survive the whole pipeline: ids 1 and 3 carry identical numbers in every run
    for run in range(1, 13):
        by = {r["Strategy"]: r for r in rows if r["Run Index"] == run}
        a, b = by[tm.DESCRIPTIONS[1]], by[tm.DESCRIPTIONS[3]]
        assert all(a[k] == b[k] for k in ("Avg Profit (€)", "Avg Drawdown (€)", "Profit/MaxDD"))


def test_error_exit_aborts_dps_like_the_reference(tmp_path):
    """A failing child makes dps.py exit with the child's code (dps.py:274-276): our CLI must exit non-zero on errors."""
    def bad(path):
        _small(path)
        text = open(path, encoding="utf-8").read().replace("avg_win: 186", "avg_win: 0")
        open(path, "w", encoding="utf-8").write(text)
    ov = dps_overlay.build_overlay(str(tmp_path / "overlay"), REF, entry=dps_overlay.ENTRY_ORACLE, config_edit=bad)
    p = dps_overlay.run_dps(ov)
    assert p.returncode != 0
    assert "exited with error code" in p.stdout
