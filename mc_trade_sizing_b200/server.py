"""Per-user daemon that keeps the CUDA context and the libmcts engines alive between CLI invocations
(orchestrator-level batching that the UNCHANGED dps.py benefits from).

`dps.py` runs the hot path as 12 short-lived `python src/trading_models.py ...` subprocesses, three at a time
(reference dps.py:211-283).  Each one pays ~1 s of NumPy import + CUDA context creation for well under a millisecond
of GPU work.  With MCTS_SERVER=1 in the environment the CLI (and run_all_strategies) becomes a thin client:

    client                                   daemon (one per user and socket)
    ------                                   ------------------------------------------------------------
    connect to the Unix socket         --->  accept; one thread per connection
      (first caller starts the daemon,       borrow an engine from the pool (one libmcts handle = its own two
       under a file lock, and waits)          CUDA streams), run mcts_simulate with the GIL released: requests
    send one JSON line (the arguments) --->   that arrive together -- dps.py keeps three in flight -- execute
    read one JSON line (20 rows)       <---   concurrently on different streams of ONE context and share the GPU
    print the reference's table              exit after MCTS_SERVER_IDLE seconds without a request (default 120)

The client imports neither NumPy nor ctypes, so a served invocation costs the interpreter's start-up plus a socket
round trip.  The socket lives in a directory only its owner can enter (mode 0700).  There is no CPU fallback here
either: if the daemon cannot create an engine (no B200, missing libmcts.so) the client raises and the CLI exits
non-zero, which is what dps.py:274-276 checks.

    python -m mc_trade_sizing_b200.server --serve [--idle 120] [--engines 3]   # run in the foreground
    python -m mc_trade_sizing_b200.server --stop | --ping | --stats
"""
from __future__ import annotations

import json
import os
import socket
import sys
import threading
import time

PKG_PARENT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROTOCOL = 1


def socket_dir() -> str:
    base = os.environ.get("MCTS_SERVER_DIR")
    if base:
        return base
    root = os.environ.get("XDG_RUNTIME_DIR") or "/tmp"
    return os.path.join(root, "mcts-%d" % os.getuid())


def socket_path() -> str:
    return os.environ.get("MCTS_SERVER_SOCKET") or os.path.join(socket_dir(), "server.sock")


def _private_dir(d: str) -> str:
    """Create the socket directory (0700) or verify that an existing one belongs to this user and nobody else can write to it:
    whoever can replace the socket could answer the CLI's requests."""
    os.makedirs(d, mode=0o700, exist_ok=True)
    st = os.stat(d)
    if st.st_uid != os.getuid() or (st.st_mode & 0o022):
        raise RuntimeError("mcts server directory %s is not private to uid %d (owner %d, mode %o)" % (d, os.getuid(), st.st_uid, st.st_mode & 0o777))
    return d


# ---- client -------------------------------------------------------------------------------------------------


def _connect(timeout=None):
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.settimeout(timeout)
    try:
        s.connect(socket_path())
    except OSError:
        s.close()
        return None
    return s


def _spawn_daemon():
    import subprocess
    d = _private_dir(socket_dir())
    env = dict(os.environ)
    env.pop("MCTS_SERVER", None)  # the daemon itself computes
    env["PYTHONPATH"] = PKG_PARENT + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
    log = open(os.path.join(d, "server.log"), "ab")
    subprocess.Popen([sys.executable, "-m", "mc_trade_sizing_b200.server", "--serve"], stdin=subprocess.DEVNULL, stdout=log, stderr=log,
                     env=env, cwd=PKG_PARENT, start_new_session=True, close_fds=True)
    log.close()


def ensure_server(wait_s: float = 120.0):
    """Connected socket to the daemon, starting it if nobody is listening.  Concurrent first callers (dps.py starts
    three CLI processes at once) serialise on a file lock so that exactly one daemon is spawned."""
    s = _connect(timeout=wait_s)
    if s is not None:
        return s
    import fcntl
    d = _private_dir(socket_dir())
    with open(os.path.join(d, "spawn.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        s = _connect(timeout=wait_s)
        if s is not None:
            return s
        try:
            os.unlink(socket_path())  # stale socket of a dead daemon
        except OSError:
            pass
        _spawn_daemon()
        deadline = time.time() + wait_s
        while time.time() < deadline:
            s = _connect(timeout=wait_s)
            if s is not None:
                return s
            time.sleep(0.02)
    log = os.path.join(d, "server.log")
    tail = ""
    try:
        with open(log, "rb") as f:
            tail = f.read()[-2000:].decode("utf-8", "replace")
    except OSError:
        pass
    raise RuntimeError("mcts server did not come up within %.0f s (there is no CPU fallback); log %s:\n%s" % (wait_s, log, tail))


def request(msg: dict, start: bool = True, timeout: float = 3600.0) -> dict:
    s = ensure_server() if start else _connect(timeout=timeout)
    if s is None:
        raise ConnectionError("no mcts server is listening on " + socket_path())
    try:
        s.settimeout(timeout)
        s.sendall((json.dumps(msg) + "\n").encode("utf-8"))
        buf = bytearray()
        while not buf.endswith(b"\n"):
            part = s.recv(1 << 16)
            if not part:
                break
            buf += part
    finally:
        s.close()
    if not buf:
        raise ConnectionError("mcts server closed the connection without a reply")
    return json.loads(buf.decode("utf-8"))


def client_run(args, kwargs, seed, device, dedup, chunk_paths):
    """run_all_strategies through the daemon: the 20 result rows (description + 9 floats) in strategy-id order.
    Errors come back with their kind so that the caller sees the exception type a direct call raises."""
    rep = request({"op": "run", "protocol": PROTOCOL, "args": list(args), "kwargs": kwargs, "seed": int(seed),
                   "device": list(device) if isinstance(device, tuple) else device, "dedup": bool(dedup), "chunk_paths": int(chunk_paths)})
    if not rep.get("ok"):
        if rep.get("kind") == "ValueError":
            raise ValueError(rep.get("error", ""))
        raise RuntimeError("mcts server: " + rep.get("error", "unknown error"))
    rows = [(r[0],) + tuple(float(x) for x in r[1:]) for r in rep["rows"]]
    if "numpy" in sys.modules:  # API users get the reference's np.float64; the bare CLI never imports NumPy
        import numpy as np
        rows = [(r[0],) + tuple(np.float64(x) for x in r[1:]) for r in rows]
    return rows


# ---- daemon -------------------------------------------------------------------------------------------------


class _Pool:
    """Engines per device spec, created on demand up to `limit` each; a request borrows one for its call."""

    def __init__(self, limit: int):
        self.limit, self.lock, self.free, self.made = limit, threading.Condition(), {}, {}

    def borrow(self, spec):
        from . import engine as _engine
        key = tuple(spec) if isinstance(spec, list) else spec
        with self.lock:
            while True:
                if self.free.get(key):
                    return key, self.free[key].pop()
                if self.made.get(key, 0) < self.limit:
                    self.made[key] = self.made.get(key, 0) + 1
                    break
                self.lock.wait()
        try:
            eng = _engine.EngineGroup(None) if key == "all" else _engine.EngineGroup(list(key)) if isinstance(key, tuple) else _engine.Engine(int(key))
        except BaseException:
            with self.lock:
                self.made[key] -= 1
                self.lock.notify()
            raise
        return key, eng

    def give_back(self, key, eng):
        with self.lock:
            self.free.setdefault(key, []).append(eng)
            self.lock.notify()


def serve(idle_s: float = 120.0, engines: int = 3) -> int:
    """Foreground server loop; returns the process exit code."""
    import socketserver
    from . import engine as _engine
    from . import trading_models as tm

    path = socket_path()
    _private_dir(os.path.dirname(path))
    pool = _Pool(engines)
    stats = {"requests": 0, "runs": 0, "errors": 0, "started": time.time(), "busy": 0, "max_concurrent": 0, "last": time.time()}
    stats_lock = threading.Lock()

    def run(msg):
        params = _engine.make_params(*msg["args"], **msg.get("kwargs", {}))
        key, eng = pool.borrow(msg.get("device", 0))
        try:
            out = eng.simulate(params, seed=int(msg["seed"]), options=eng.options(chunk_paths=int(msg.get("chunk_paths", 0)), dedup=bool(msg.get("dedup"))))
        finally:
            pool.give_back(key, eng)
        return [[tm.DESCRIPTIONS[i + 1]] + [float(x) for x in out["summary"][i]] for i in range(20)]

    class Handler(socketserver.StreamRequestHandler):
        def handle(self):
            line = self.rfile.readline()
            if not line:
                return
            with stats_lock:
                stats["requests"] += 1
                stats["busy"] += 1
                stats["max_concurrent"] = max(stats["max_concurrent"], stats["busy"])
            try:
                msg = json.loads(line.decode("utf-8"))
                op = msg.get("op")
                if op == "run" and msg.get("protocol", PROTOCOL) != PROTOCOL:  # a daemon left over from another install
                    rep = {"ok": False, "kind": "RuntimeError", "error": "protocol %r, this daemon speaks %d: stop it with "
                           "`python -m mc_trade_sizing_b200.server --stop`" % (msg.get("protocol"), PROTOCOL)}
                elif op == "run":
                    rep = {"ok": True, "rows": run(msg)}
                    with stats_lock:
                        stats["runs"] += 1
                elif op == "ping":
                    rep = {"ok": True, "pid": os.getpid(), "protocol": PROTOCOL}
                elif op == "stats":
                    with stats_lock:
                        rep = {"ok": True, "pid": os.getpid(), **{k: v for k, v in stats.items()}}
                elif op == "shutdown":
                    rep = {"ok": True, "pid": os.getpid()}
                    threading.Thread(target=server.shutdown, daemon=True).start()
                else:
                    rep = {"ok": False, "kind": "ValueError", "error": "unknown op %r" % (op,)}
            except ValueError as e:  # bad probabilities, zero paths: what NumPy raises in the reference
                rep = {"ok": False, "kind": "ValueError", "error": str(e)}
            except Exception as e:  # MctsError (no device, unsupported size ...): the client raises RuntimeError
                rep = {"ok": False, "kind": type(e).__name__, "error": str(e)}
            if not rep.get("ok"):
                with stats_lock:
                    stats["errors"] += 1
            try:
                self.wfile.write((json.dumps(rep) + "\n").encode("utf-8"))
            except OSError:
                pass
            with stats_lock:
                stats["busy"] -= 1
                stats["last"] = time.time()

    class Server(socketserver.ThreadingMixIn, socketserver.UnixStreamServer):
        daemon_threads = True
        allow_reuse_address = True
        request_queue_size = 64

    try:
        os.unlink(path)
    except OSError:
        pass
    # fail before listening if this process cannot compute at all: the client then reports the reason from the log
    key, eng = pool.borrow(int(os.environ.get("MCTS_DEVICE", "0")))
    pool.give_back(key, eng)
    old_umask = os.umask(0o077)
    try:
        server = Server(path, Handler)
    finally:
        os.umask(old_umask)

    def watchdog():
        while True:
            time.sleep(min(1.0, max(idle_s / 4, 0.05)))
            with stats_lock:
                idle = stats["busy"] == 0 and time.time() - stats["last"] > idle_s
            if idle:
                server.shutdown()
                return

    threading.Thread(target=watchdog, daemon=True).start()
    print("[mcts server] pid %d listening on %s (idle timeout %.0f s, %d engines per device)" % (os.getpid(), path, idle_s, engines), flush=True)
    try:
        server.serve_forever(poll_interval=0.05)
    finally:
        server.server_close()
        try:
            os.unlink(path)
        except OSError:
            pass
    print("[mcts server] pid %d exiting after %d runs" % (os.getpid(), stats["runs"]), flush=True)
    return 0


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="mc_trade_sizing_b200 engine daemon")
    ap.add_argument("--serve", action="store_true")
    ap.add_argument("--idle", type=float, default=float(os.environ.get("MCTS_SERVER_IDLE", "120")))
    ap.add_argument("--engines", type=int, default=int(os.environ.get("MCTS_SERVER_ENGINES", "3")))
    ap.add_argument("--stop", action="store_true")
    ap.add_argument("--ping", action="store_true")
    ap.add_argument("--stats", action="store_true")
    a = ap.parse_args(argv)
    if a.serve:
        return serve(a.idle, a.engines)
    op = "shutdown" if a.stop else "stats" if a.stats else "ping"
    try:
        print(json.dumps(request({"op": op}, start=False, timeout=10.0)))
    except (ConnectionError, OSError) as e:
        print(json.dumps({"ok": False, "error": str(e)}))
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
