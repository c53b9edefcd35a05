"""Helpers to run the reference's UNCHANGED dps.py against the drop-in: an overlay directory
holding a symlink to the reference's dps.py, a copy of its dps_config.yaml and our own src/trading_models.py.  dps.py
resolves `script_dir` from the symlink's location (dps.py:169, :219 -> our CLI, our yaml, results/ in the overlay) while
Python puts the symlink's REAL directory on sys.path (dps.py:50-56 -> the reference's own sinks).  TEST ONLY."""
import glob
import json
import os
import shutil
import sqlite3
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "tests", "stubs")


def reference_dir():
    """The reference checkout: /root/reference in the build container, the shipped copy baseline/_ref on the GPU box."""
    for d in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.exists(os.path.join(d, "dps.py")) and os.path.exists(os.path.join(d, "src", "output_handler.py")):
            return d
    return None


ENTRY_PRODUCT = '''import sys
sys.path.insert(0, %(root)r)
from mc_trade_sizing_b200.trading_models import main
main()
'''

# CPU-only variant (no GPU in the build container): the same CLI, but the hot function is answered by the test oracle,
# so that everything around it -- our stdout, dps.py's scraper, dedupe and sinks -- is exercised without a device.
ENTRY_ORACLE = '''import sys
sys.path.insert(0, %(root)r)
from mc_trade_sizing_b200 import trading_models as tm
from oracle import oracle
def run_all_strategies(*args, **kw):
    seed = tm.resolve_seed(None) & 0x7fffffff
    return tm.summary_to_rows(oracle.run_all(oracle.make_params(*args, **kw), seed)["summary"])
tm.run_all_strategies = run_all_strategies
tm.main()
'''


def build_overlay(dst, ref, entry=ENTRY_PRODUCT, config_edit=None):
    os.makedirs(os.path.join(dst, "src"), exist_ok=True)
    os.symlink(os.path.join(ref, "dps.py"), os.path.join(dst, "dps.py"))
    shutil.copyfile(os.path.join(ref, "dps_config.yaml"), os.path.join(dst, "dps_config.yaml"))
    if config_edit:
        config_edit(os.path.join(dst, "dps_config.yaml"))
    with open(os.path.join(dst, "src", "trading_models.py"), "w") as f:
        f.write(entry % {"root": ROOT})
    return dst


def run_dps(overlay, extra_env=None, timeout=900):
    """`python dps.py` with cwd = overlay (src/influx_handler.py:5 reads the yaml relative to the CWD), answering 'n' to both
    prompts (dps.py:429-431).  The stubs stand in for openpyxl / influxdb_client / flask, which this image lacks."""
    env = dict(os.environ)
    env["PYTHONPATH"] = STUBS + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
    env.update(extra_env or {})
    return subprocess.run([sys.executable, "dps.py"], cwd=overlay, input="n\nn\n", capture_output=True, text=True, env=env, timeout=timeout)


def check_outputs(overlay, expect_runs=12):
    """What dps.py:339-424 must have produced from our stdout: 20 scraped rows per run in the CSV / JSON / Parquet /
    SQLite sinks, the HTML overview and the (stub-written) workbook."""
    import pandas as pd
    res = os.path.join(overlay, "results")
    one = lambda pat: (glob.glob(os.path.join(res, pat)) or [None])[0]
    js, csv, pq, db, html, xlsx = (one(p) for p in ("simulation_runs_*.json", "simulation_runs_*.csv", "simulation_runs_*.parquet",
                                                     "simulation_results_*.db", "simulation_runs_*.html", "simulation_runs_*.xlsx"))
    for f in (js, csv, pq, db, html, xlsx):
        assert f and os.path.getsize(f) > 0, "missing sink output in %s: %s" % (res, os.listdir(res))
    rows = json.load(open(js, encoding="utf-8"))
    assert len(rows) == 20 * expect_runs, len(rows)
    assert sorted({r["Run Index"] for r in rows}) == list(range(1, expect_runs + 1))
    df = pd.read_csv(csv, sep=";", encoding="utf-8-sig")
    assert len(df) == 20 * expect_runs and len(pd.read_parquet(pq)) == 20 * expect_runs
    con = sqlite3.connect(db)
    try:
        assert con.execute("select count(*) from trade_results").fetchone()[0] == 20 * expect_runs
    finally:
        con.close()
    return rows
