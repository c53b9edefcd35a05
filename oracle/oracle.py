"""ctypes front-end of the CPU oracle (oracle/mcts_oracle.c, oracle/devgen_ref.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(mc_trade_sizing_b200) never imports this module and has no CPU fallback.

The oracle restates /root/reference/src/trading_models.py:43-168,170-404,445-522
(generators, shuffle, the 20 sizing closures, drawdown, NumPy summation order,
aggregates).  It is pinned against fixtures generated from the live reference
(tests/golden/make_golden.py, checked by tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MAX_REGIMES = 16
NSTRAT = 20
MODES = {"dynamic": 0, "markov1": 1, "markov2": 2, "regime": 3}
SUMMARY_COLS = (
    "avg_profit", "avg_drawdown", "ratio", "min_profit", "max_profit",
    "min_dd", "max_dd", "avg_per_trade", "ratio_max_dd",
)


class Regime(C.Structure):
    _fields_ = [("length", C.c_int), ("hit_rate", C.c_double), ("avg_win", C.c_double), ("avg_loss", C.c_double)]


class Params(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("num_trades", C.c_int), ("num_simulations", C.c_int), ("num_mc_shuffles", C.c_int),
        ("hit_rate", C.c_double), ("avg_win", C.c_double), ("avg_loss", C.c_double),
        ("p_win_after_win", C.c_double), ("p_win_after_loss", C.c_double),
        ("p_win_ww", C.c_double), ("p_win_wl", C.c_double), ("p_win_lw", C.c_double), ("p_win_ll", C.c_double),
        ("n_regimes", C.c_int), ("regimes", Regime * MAX_REGIMES),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, no fast-math)."""
    srcs = [os.path.join(_HERE, f) for f in ("mcts_oracle.c", "devgen_ref.c", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.oracle_sequence_length.restype = C.c_int64
        L.oracle_sequence_length.argtypes = [C.POINTER(Params)]
        L.oracle_run_all.restype = C.c_int
        L.oracle_run_all.argtypes = [C.POINTER(Params), C.c_uint32, dp, dp, dp, dp]
        L.oracle_eval_paths.restype = C.c_int
        L.oracle_eval_paths.argtypes = [dp, C.c_int64, C.c_int64, C.c_int, dp, dp, dp]
        L.oracle_equity_curves.restype = C.c_int
        L.oracle_equity_curves.argtypes = [dp, C.c_int64, C.c_int64, dp]
        L.oracle_summarise.restype = C.c_int
        L.oracle_summarise.argtypes = [dp, dp, C.c_int64, C.c_int, dp]
        L.oracle_np_sum.restype = C.c_double
        L.oracle_np_sum.argtypes = [dp, C.c_int64]
        L.oracle_rng_taps.restype = None
        L.oracle_rng_taps.argtypes = [C.c_uint32, C.c_int, dp, C.c_int, dp]
        L.devgen_sequences.restype = C.c_int64
        L.devgen_sequences.argtypes = [C.POINTER(Params), C.c_uint64, C.c_int64, C.c_int64, dp]
        L.devgen_philox.restype = None
        L.devgen_philox.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_bench.restype = C.c_double
        L.oracle_bench.argtypes = [C.POINTER(Params), C.c_uint32, C.c_int, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def make_params(hit_rate, avg_win, avg_loss, num_trades, num_simulations, num_mc_shuffles,
                use_markov=False, p_win_after_win=0.7, p_win_after_loss=0.5,
                use_markov2=False, p_win_ww=0.8, p_win_wl=0.6, p_win_lw=0.5, p_win_ll=0.3,
                use_regime=False, regimes=None) -> Params:
    """Same argument list as run_all_strategies (trading_models.py:445-450);
    mode precedence regime > markov2 > markov1 > dynamic as in :477-490."""
    p = Params()
    p.mode = 3 if use_regime else 2 if use_markov2 else 1 if use_markov else 0
    p.num_trades, p.num_simulations, p.num_mc_shuffles = int(num_trades), int(num_simulations), int(num_mc_shuffles)
    p.hit_rate, p.avg_win, p.avg_loss = float(hit_rate), float(avg_win), float(avg_loss)
    p.p_win_after_win, p.p_win_after_loss = float(p_win_after_win), float(p_win_after_loss)
    p.p_win_ww, p.p_win_wl, p.p_win_lw, p.p_win_ll = map(float, (p_win_ww, p_win_wl, p_win_lw, p_win_ll))
    if regimes is None:
        p.n_regimes = -1
    else:
        if len(regimes) > MAX_REGIMES:
            raise ValueError("oracle supports at most %d regimes" % MAX_REGIMES)
        p.n_regimes = len(regimes)
        for i, r in enumerate(regimes):
            p.regimes[i] = Regime(int(r["length"]), float(r["hit_rate"]), float(r["avg_win"]), float(r["avg_loss"]))
    return p


def _check(rc):
    if rc == -1:
        raise ValueError("oracle: invalid probabilities (NumPy's choice() would raise)")
    if rc == -2:
        raise ValueError("oracle: empty sequence or zero paths (NumPy's min() would raise)")
    if rc != 0:
        raise RuntimeError("oracle error %d" % rc)


def sequence_length(params: Params) -> int:
    return int(lib().oracle_sequence_length(C.byref(params)))


def run_all(params: Params, seed: int, per_path: bool = False, sequences: bool = False):
    """run_all_strategies after np.random.seed(seed).  Returns a dict with
    'summary' [20,9] (strategy-id order) and optionally 'profit'/'dd' [20,P]
    and 'seqs' [P,L]."""
    P = params.num_simulations * params.num_mc_shuffles
    Lseq = sequence_length(params)
    out = {"summary": np.zeros((NSTRAT, 9))}
    profit = dd = seqs = None
    if per_path:
        profit, dd = np.zeros((NSTRAT, max(P, 0))), np.zeros((NSTRAT, max(P, 0)))
    if sequences:
        seqs = np.zeros((max(P, 0), max(Lseq, 0)))
    _check(lib().oracle_run_all(C.byref(params), C.c_uint32(seed & 0xFFFFFFFF), _dp(out["summary"]),
                                _dp(profit), _dp(dd), _dp(seqs)))
    if per_path:
        out["profit"], out["dd"] = profit, dd
    if sequences:
        out["seqs"] = seqs
    return out


def eval_paths(values: np.ndarray, num_trades: int | None = None):
    """Evaluate all 20 strategies on caller-supplied sequences values[P,T]."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    P, T = v.shape
    profit, dd, summary = np.zeros((NSTRAT, P)), np.zeros((NSTRAT, P)), np.zeros((NSTRAT, 9))
    _check(lib().oracle_eval_paths(_dp(v), P, T, int(T if num_trades is None else num_trades),
                                   _dp(profit), _dp(dd), _dp(summary)))
    return {"profit": profit, "dd": dd, "summary": summary}


def equity_curves(values: np.ndarray) -> np.ndarray:
    """[20, P, T] equity curves (sequential cumsum of contributions) of caller-supplied sequences."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    P, T = v.shape
    curves = np.zeros((NSTRAT, P, T))
    _check(lib().oracle_equity_curves(_dp(v), P, T, _dp(curves)))
    return curves


def summarise(profit: np.ndarray, dd: np.ndarray, num_trades: int) -> np.ndarray:
    profit = np.ascontiguousarray(profit, dtype=np.float64)
    dd = np.ascontiguousarray(dd, dtype=np.float64)
    summary = np.zeros((NSTRAT, 9))
    _check(lib().oracle_summarise(_dp(profit), _dp(dd), profit.shape[1], int(num_trades), _dp(summary)))
    return summary


def np_sum(a: np.ndarray) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().oracle_np_sum(_dp(a), a.size))


def rng_taps(seed: int, n_doubles: int, shuffle_n: int):
    d, s = np.zeros(n_doubles), np.zeros(shuffle_n)
    lib().oracle_rng_taps(C.c_uint32(seed), n_doubles, _dp(d), shuffle_n, _dp(s))
    return d, s


def bench(params: Params, seed0: int, n_threads: int) -> float:
    """Wall seconds for n_threads concurrent, independently seeded runs of `params`."""
    sink = C.c_double()
    secs = float(lib().oracle_bench(C.byref(params), C.c_uint32(seed0), int(n_threads), C.byref(sink)))
    if secs < 0:
        raise RuntimeError("oracle bench failed")
    return secs


def devgen_sequences(params: Params, seed: int, first_path: int = 0, n: int | None = None) -> np.ndarray:
    """Independent CPU restatement (oracle/devgen_ref.c) of the DEVICE generator specification:
    the post-shuffle sequences mcts_simulate evaluates for paths [first_path, first_path+n)."""
    P = params.num_simulations * params.num_mc_shuffles
    n = P - first_path if n is None else n
    Lseq = sequence_length(params)
    out = np.zeros((max(n, 0), max(Lseq, 0)))
    rc = lib().devgen_sequences(C.byref(params), C.c_uint64(seed & (2**64 - 1)), int(first_path), int(n), _dp(out))
    if rc < 0:
        raise ValueError("devgen: invalid probabilities (NumPy's choice() would raise)")
    assert rc == Lseq or n == 0, (rc, Lseq)
    return out


def devgen_philox(key, ctr):
    c = (C.c_uint32 * 4)(*ctr)
    o = (C.c_uint32 * 4)()
    lib().devgen_philox(key[0], key[1], c, o)
    return list(o)
