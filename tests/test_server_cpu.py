"""The engine daemon's protocol and life cycle on CPU.  The daemon runs in a thread of the test
process with the engine class replaced by the CPU harness of the device code (TEST ONLY), so everything but the GPU
call is the product's: socket set-up, the thin client in run_all_strategies / the CLI, concurrent requests on pooled
engines, error propagation with the right exception type, stats, idle shutdown.  The real daemon on a B200 is covered
by tests/test_gpu_parity.py::test_server_daemon_serves_cli."""
import os
import subprocess
import sys
import threading
import time

import numpy as np
import pytest

import harness
from conftest import assert_bitexact
from mc_trade_sizing_b200 import engine as engine_mod
from mc_trade_sizing_b200 import server
from mc_trade_sizing_b200 import trading_models as tm
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeEngine:
    """Same simulate() contract as Engine, computed by the g++ build of the device headers."""
    created = 0
    options = staticmethod(engine_mod.Engine.options)

    def __init__(self, device=0):
        type(self).created += 1
        self.device = device

    def simulate(self, params, seed=0, per_path=False, options=None):
        if params.num_simulations * params.num_mc_shuffles <= 0:
            raise ValueError("zero paths (np.min of an empty list raises in the reference)")
        time.sleep(0.05)  # lets concurrent requests overlap
        out = harness.simulate(params, seed, dedup=bool(options.dedup) if options is not None else False)
        return {"summary": oracle.summarise(out["profit"], out["dd"], params.num_trades)}


@pytest.fixture()
def daemon(tmp_path, monkeypatch):
    sock = str(tmp_path / "d" / "server.sock")
    monkeypatch.setenv("MCTS_SERVER_SOCKET", sock)
    monkeypatch.setenv("MCTS_SERVER_DIR", str(tmp_path / "d"))
    monkeypatch.setattr(engine_mod, "Engine", FakeEngine)
    FakeEngine.created = 0
    t = threading.Thread(target=server.serve, kwargs=dict(idle_s=1.0, engines=3), daemon=True)
    t.start()
    for _ in range(500):
        if os.path.exists(sock):
            break
        time.sleep(0.01)
    yield t
    if t.is_alive():
        try:
            server.request({"op": "shutdown"}, start=False, timeout=5)
        except (ConnectionError, OSError):
            pass
        t.join(timeout=10)


ARGS = (0.82, 186.0, 219.0, 120, 6, 5)


def _expected(seed, **kw):
    prm = engine_mod.make_params(*ARGS, **kw)
    out = harness.simulate(prm, seed)
    return tm.summary_to_rows(oracle.summarise(out["profit"], out["dd"], ARGS[3]))


def test_served_call_equals_the_direct_result(daemon, monkeypatch):
    monkeypatch.setenv("MCTS_SERVER", "1")
    got = tm.run_all_strategies(*ARGS, seed=31, use_markov=True, p_win_after_win=0.75, p_win_after_loss=0.6)
    exp = _expected(31, use_markov=True, p_win_after_win=0.75, p_win_after_loss=0.6)
    assert [r[0] for r in got] == [r[0] for r in exp]
    assert_bitexact(np.array([r[1:] for r in got], dtype=np.float64), np.array([r[1:] for r in exp], dtype=np.float64), "served rows")
    assert isinstance(got[0][1], np.float64)            # the reference's return type, once NumPy is loaded
    assert server.request({"op": "ping"}, start=False)["pid"] == os.getpid()


def test_concurrent_clients_share_pooled_engines(daemon, monkeypatch):
    monkeypatch.setenv("MCTS_SERVER", "1")
    results, errors = {}, []

    def client(i):
        try:
            results[i] = tm.run_all_strategies(*ARGS, seed=100 + i)
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=client, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not errors
    for i in range(6):
        exp = _expected(100 + i)
        assert_bitexact(np.array([r[1:] for r in results[i]], dtype=np.float64), np.array([r[1:] for r in exp], dtype=np.float64), "client %d" % i)
    st = server.request({"op": "stats"}, start=False)
    assert st["runs"] == 6 and st["errors"] == 0
    assert 2 <= st["max_concurrent"] and FakeEngine.created <= 3 + 1      # three pooled engines (+ the start-up probe's, reused)


def test_errors_keep_their_type(daemon, monkeypatch):
    monkeypatch.setenv("MCTS_SERVER", "1")
    with pytest.raises(ValueError):
        tm.run_all_strategies(0.82, 186.0, 219.0, 100, 0, 5, seed=1)     # zero paths: ValueError like NumPy's
    assert server.request({"op": "stats"}, start=False)["errors"] == 1


def test_other_protocol_version_is_refused(daemon):
    rep = server.request({"op": "run", "protocol": server.PROTOCOL + 1, "args": list(ARGS), "kwargs": {}, "seed": 1}, start=False)
    assert rep["ok"] is False and rep["kind"] == "RuntimeError" and "protocol" in rep["error"]
    assert server.request({"op": "stats"}, start=False)["runs"] == 0


def test_cli_client_is_thin_and_prints_the_reference_table(daemon):
    """The CLI through the daemon: same stdout as a direct run, and the client process never loads NumPy or ctypes."""
    code = ("import sys; sys.argv = ['trading_models.py', '--hit_rate', '0.82', '--avg_win', '186', '--avg_loss', '219', '--num_simulations', '6',"
            " '--num_trades', '120', '--num_mc_shuffles', '5']; sys.path.insert(0, %r);"
            " from mc_trade_sizing_b200.trading_models import main; main();"
            " assert 'numpy' not in sys.modules and 'ctypes' not in sys.modules, sorted(sys.modules)") % ROOT
    env = dict(os.environ, MCTS_SERVER="1", MCTS_SEED="31")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=60)
    assert p.returncode == 0, p.stderr
    rows = _expected(31)
    lines = [ln for ln in p.stdout.split("\n") if ln.startswith(tuple(tm.DESCRIPTIONS.values()))]
    assert lines[:20] == [tm.format_row(r) for r in rows]


def test_idle_daemon_exits(daemon):
    assert server.request({"op": "ping"}, start=False)["ok"]
    daemon.join(timeout=15)          # idle_s = 1.0
    assert not daemon.is_alive()
    assert not os.path.exists(os.environ["MCTS_SERVER_SOCKET"])
    with pytest.raises(ConnectionError):
        server.request({"op": "ping"}, start=False, timeout=2)
