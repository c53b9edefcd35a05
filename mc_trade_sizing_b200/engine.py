"""Engine: one libmcts handle (one B200) behind a small Python object.  Host-side mirror of the
reference's call structure: parameters in, the result table out; all compute is in libmcts.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import (MODE_DYNAMIC, MODE_MARKOV1, MODE_MARKOV2, MODE_REGIME, NSTRAT, Agg, ChunkResult, MctsError, Options,
                   Params, Regime, Timing, check, lib)

DEFAULT_SEED = 12345

SUMMARY_COLS = ("avg_profit", "avg_drawdown", "ratio", "min_profit", "max_profit", "min_dd", "max_dd",
                "avg_per_trade", "ratio_max_dd")


def make_params(hit_rate, avg_win, avg_loss, num_trades, num_simulations, num_mc_shuffles,
                use_markov=False, p_win_after_win=0.7, p_win_after_loss=0.5,
                use_markov2=False, p_win_ww=0.8, p_win_wl=0.6, p_win_lw=0.5, p_win_ll=0.3,
                use_regime=False, regimes=None) -> Params:
    """Flatten run_all_strategies' argument list (reference src/trading_models.py:445-450); generator
    precedence regime > markov2 > markov1 > dynamic as in :477-490."""
    p = Params()
    p.mode = MODE_REGIME if use_regime else MODE_MARKOV2 if use_markov2 else MODE_MARKOV1 if use_markov else MODE_DYNAMIC
    p.num_trades, p.num_simulations, p.num_mc_shuffles = int(num_trades), int(num_simulations), int(num_mc_shuffles)
    p.hit_rate, p.avg_win, p.avg_loss = float(hit_rate), float(avg_win), float(avg_loss)
    p.p_win_after_win, p.p_win_after_loss = float(p_win_after_win), float(p_win_after_loss)
    p.p_win_ww, p.p_win_wl, p.p_win_lw, p.p_win_ll = map(float, (p_win_ww, p_win_wl, p_win_lw, p_win_ll))
    if regimes is None or not use_regime:
        p.n_regimes = -1
    else:
        if len(regimes) > _ffi.MAX_REGIMES:
            raise MctsError(_ffi.EUNSUPPORTED, "at most %d regimes are supported" % _ffi.MAX_REGIMES)
        p.n_regimes = len(regimes)
        for i, r in enumerate(regimes):
            p.regimes[i] = Regime(int(r["length"]), float(r["hit_rate"]), float(r["avg_win"]), float(r["avg_loss"]))
    return p


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _raise(rc):
    """Map libmcts status codes to the exception types the reference would raise."""
    if rc >= 0:
        return
    msg = lib().mcts_last_error()
    msg = msg.decode("utf-8", "replace") if msg else ""
    if rc in (_ffi.EINVAL, _ffi.EEMPTY):
        raise ValueError(msg)  # NumPy raises ValueError for bad probabilities / empty reductions
    raise MctsError(rc, msg)


def agg_to_rows(agg, num_trades: int) -> np.ndarray:
    rows = np.zeros((NSTRAT, 9))
    _raise(lib().mcts_summary_rows(agg, int(num_trades), _dp(rows)))
    return rows


class Engine:
    """One device.  Not thread-safe; create one per thread / process / GPU."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().mcts_create(int(device), C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().mcts_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def options(chunk_paths=0, dedup=False, int_fast_path=True, block_threads=0, serial_chunks=False) -> Options:
        return Options(int(chunk_paths), int(bool(dedup)), int(bool(int_fast_path)), int(block_threads), int(bool(serial_chunks)))

    # -- on-device RNG ------------------------------------------------------------------------
    def simulate(self, params: Params, seed: int = DEFAULT_SEED, per_path: bool = False, options: Options | None = None):
        """Returns {'agg': Agg[20], 'summary': [20,9] (strategy-id order)} and, with per_path,
        'profit' / 'dd' [20,P]."""
        P = params.num_simulations * params.num_mc_shuffles
        agg = (Agg * NSTRAT)()
        profit = dd = None
        if per_path and P > 0:
            profit, dd = np.empty((NSTRAT, P)), np.empty((NSTRAT, P))
        opt = options if options is not None else self.options()
        _raise(lib().mcts_simulate(self._h, C.byref(params), C.c_uint64(seed & (2**64 - 1)), C.byref(opt), agg, _dp(profit), _dp(dd)))
        out = {"agg": agg, "summary": agg_to_rows(agg, params.num_trades)}
        if per_path:
            out["profit"], out["dd"] = profit, dd
        return out

    def sequences(self, params: Params, seed: int = DEFAULT_SEED, first_path: int = 0, n: int | None = None) -> np.ndarray:
        """The post-shuffle sequences the device generator uses for paths [first_path, first_path+n)."""
        P = params.num_simulations * params.num_mc_shuffles
        n = P - first_path if n is None else n
        L = int(lib().mcts_sequence_length(C.byref(params)))
        if L < 0:
            _raise(L)
        vals = np.empty((max(n, 0), max(L, 0)))
        _raise(lib().mcts_sequences(self._h, C.byref(params), C.c_uint64(seed & (2**64 - 1)), int(first_path), int(n), _dp(vals)))
        return vals

    # -- replay -------------------------------------------------------------------------------
    def replay(self, values, num_trades: int | None = None, per_path: bool = True, options: Options | None = None):
        """Evaluate caller-supplied post-shuffle sequences values[P,T] (any real dtype; the reference's
        int64 regime arrays are exact in fp64 below 2^53)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.ndim != 2:
            raise ValueError("values must be [paths, trades]")
        P, T = v.shape
        agg = (Agg * NSTRAT)()
        profit = dd = None
        if per_path and P > 0:
            profit, dd = np.empty((NSTRAT, P)), np.empty((NSTRAT, P))
        opt = options if options is not None else self.options()
        _raise(lib().mcts_replay(self._h, _dp(v), P, T, C.byref(opt), agg, _dp(profit), _dp(dd)))
        out = {"agg": agg, "summary": agg_to_rows(agg, T if num_trades is None else num_trades)}
        if per_path:
            out["profit"], out["dd"] = profit, dd
        return out

    # -- multi-device building blocks -----------------------------------------------------------
    @staticmethod
    def auto_chunk(P: int, n_shards: int) -> int:
        """Chunk size for a plan that is split over n_shards devices / ranks (mcts_auto_chunk)."""
        return int(lib().mcts_auto_chunk(int(P), int(n_shards)))

    @staticmethod
    def plan_shards(P: int, chunk_paths: int, n_shards: int):
        """Chunk ranges [lo, hi) per shard, contiguous and balanced by path count (mcts_plan_shards)."""
        b = np.zeros(n_shards + 1, dtype=np.int64)
        _raise(lib().mcts_plan_shards(int(P), int(chunk_paths), int(n_shards), b.ctypes.data_as(C.POINTER(C.c_int64))))
        return [(int(b[r]), int(b[r + 1])) for r in range(n_shards)]

    @staticmethod
    def plan(P: int, chunk_paths: int = 0):
        n = int(lib().mcts_plan_count(int(P), int(chunk_paths)))
        first, count = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        if n:
            _raise(lib().mcts_plan(int(P), int(chunk_paths), first.ctypes.data_as(C.POINTER(C.c_int64)),
                                   count.ctypes.data_as(C.POINTER(C.c_int64))))
        return first, count

    def simulate_chunks(self, params: Params, seed: int, chunk_begin: int, chunk_end: int, options: Options | None = None):
        n = max(chunk_end - chunk_begin, 0)
        recs = (ChunkResult * max(n, 1))()
        opt = options if options is not None else self.options()
        _raise(lib().mcts_simulate_chunks(self._h, C.byref(params), C.c_uint64(seed & (2**64 - 1)), C.byref(opt),
                                          int(chunk_begin), int(chunk_end), recs))
        return recs, n

    @staticmethod
    def combine(P: int, chunk_paths: int, recs, n: int):
        agg = (Agg * NSTRAT)()
        _raise(lib().mcts_combine(int(P), int(chunk_paths), recs, int(n), agg))
        return agg

    # -- equity-curve export (extension) ------------------------------------------------------
    def export_curves(self, params: Params, seed: int, first_path: int, n: int, to_host: bool = True, dev=None):
        """Equity curves [20, ceil(L/2), n, 2] of paths [first_path, first_path+n).  With to_host=False
        only the device write is performed (bandwidth measurement / streaming) and None is returned.
        `dev` is an optional device buffer from device_alloc() that is reused instead of a fresh one."""
        L = int(lib().mcts_sequence_length(C.byref(params)))
        half = (L + 1) // 2
        nbytes = NSTRAT * half * n * 2 * 8
        own = dev is None
        if own:
            dev = self.device_alloc(nbytes)
        try:
            _raise(lib().mcts_export_curves(self._h, C.byref(params), C.c_uint64(seed & (2**64 - 1)), int(first_path), int(n), dev))
            if not to_host:
                return None
            out = np.empty((NSTRAT, half, n, 2))
            _raise(lib().mcts_copy_to_host(self._h, out.ctypes.data_as(C.c_void_p), dev, nbytes))
            return out
        finally:
            if own:
                self.device_free(dev)

    def device_alloc(self, nbytes: int):
        dev = lib().mcts_device_alloc(self._h, int(nbytes))
        if not dev:
            _raise(_ffi.ENOMEM)
        return dev

    def device_free(self, dev):
        lib().mcts_device_free(self._h, dev)

    # -- measurement ----------------------------------------------------------------------------
    def timing(self) -> dict:
        t = Timing()
        _raise(lib().mcts_get_timing(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timing._fields_}

    def stream_ptr(self) -> int:
        """cudaStream_t of this engine as an integer (wrap with torch.cuda.ExternalStream for event timing)."""
        return int(lib().mcts_get_stream(self._h) or 0)

    def measure_peaks(self) -> dict:
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _raise(lib().mcts_measure_peaks(self._h, C.byref(a), C.byref(b), C.byref(c)))
        d = C.c_double()
        _raise(lib().mcts_measure_step_ceiling(self._h, C.byref(d)))
        e, f = C.c_double(), C.c_double()
        _raise(lib().mcts_measure_mix_ceiling(self._h, C.byref(e), C.byref(f)))
        return {"fp64_inst_per_s": a.value, "int_inst_per_s": b.value, "sm_clock_mhz": c.value, "step_ceiling_per_s": d.value,
                "mix_ceiling_per_s": e.value, "mix_ceiling_ipc": f.value}


class EngineGroup:
    """Several B200s of one box behind one call (mcts_group_*): same simulate() contract and bit-identical results
    as one Engine; the path range is cut into tree-node chunks, each device evaluates a contiguous run on its own
    host thread inside libmcts, the chunk records are folded on the host in NumPy's order."""

    def __init__(self, devices=None):
        self._g = C.c_void_p()
        if devices is None:
            check(lib().mcts_group_create(None, 0, C.byref(self._g)))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            check(lib().mcts_group_create(arr, len(devices), C.byref(self._g)))
        self.devices = [int(lib().mcts_group_device(self._g, i)) for i in range(int(lib().mcts_group_size(self._g)))]

    def close(self):
        if getattr(self, "_g", None) is not None and self._g.value:
            lib().mcts_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    options = staticmethod(Engine.options)

    def simulate(self, params: Params, seed: int = DEFAULT_SEED, per_path: bool = False, options: Options | None = None):
        P = params.num_simulations * params.num_mc_shuffles
        agg = (Agg * NSTRAT)()
        profit = dd = None
        if per_path and P > 0:
            profit, dd = np.empty((NSTRAT, P)), np.empty((NSTRAT, P))
        opt = options if options is not None else self.options()
        _raise(lib().mcts_group_simulate(self._g, C.byref(params), C.c_uint64(seed & (2**64 - 1)), C.byref(opt), agg, _dp(profit), _dp(dd)))
        out = {"agg": agg, "summary": agg_to_rows(agg, params.num_trades)}
        if per_path:
            out["profit"], out["dd"] = profit, dd
        return out

    def timing(self) -> list:
        """Per-device timing of the last simulate()."""
        t = (Timing * max(len(self.devices), 1))()
        _raise(lib().mcts_group_get_timing(self._g, t))
        return [{k: getattr(t[i], k) for k, _ in Timing._fields_} for i in range(len(self.devices))]
