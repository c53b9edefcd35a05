"""ctypes binding of libmcts.so (include/mcts.h).  This is the stub a maintainer of the reference
adds next to src/trading_models.py.  There is no CPU fallback: if the shared
object is missing or no B200 is visible, importing succeeds but every call raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCTS_LIB") or os.path.join(HERE, "libmcts.so")

NSTRAT = 20
MAX_REGIMES = 16
ABI_VERSION = 2

MODE_DYNAMIC, MODE_MARKOV1, MODE_MARKOV2, MODE_REGIME = 0, 1, 2, 3

OK, EINVAL, EEMPTY, ENOMEM, ECUDA, ENODEVICE, EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6


class Regime(C.Structure):
    _fields_ = [("length", C.c_int32), ("hit_rate", C.c_double), ("avg_win", C.c_double), ("avg_loss", C.c_double)]


class Params(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("num_trades", C.c_int32), ("num_simulations", C.c_int32), ("num_mc_shuffles", C.c_int32),
        ("hit_rate", C.c_double), ("avg_win", C.c_double), ("avg_loss", C.c_double),
        ("p_win_after_win", C.c_double), ("p_win_after_loss", C.c_double),
        ("p_win_ww", C.c_double), ("p_win_wl", C.c_double), ("p_win_lw", C.c_double), ("p_win_ll", C.c_double),
        ("n_regimes", C.c_int32), ("regimes", Regime * MAX_REGIMES),
    ]


class Agg(C.Structure):
    _fields_ = [("sum_profit", C.c_double), ("sum_dd", C.c_double), ("min_profit", C.c_double), ("max_profit", C.c_double),
                ("min_dd", C.c_double), ("max_dd", C.c_double), ("count", C.c_int64)]


class Options(C.Structure):
    _fields_ = [("chunk_paths", C.c_int64), ("dedup", C.c_int32), ("int_fast_path", C.c_int32),
                ("block_threads", C.c_int32), ("serial_chunks", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("gen_ms", C.c_double), ("paths_ms", C.c_double), ("reduce_ms", C.c_double),
                ("launches", C.c_int64), ("paths_launches", C.c_int64), ("trade_steps", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


class ChunkResult(C.Structure):
    _fields_ = [("first_path", C.c_int64), ("n_paths", C.c_int64),
                ("sum_profit", C.c_double * NSTRAT), ("sum_dd", C.c_double * NSTRAT),
                ("min_profit", C.c_double * NSTRAT), ("max_profit", C.c_double * NSTRAT),
                ("min_dd", C.c_double * NSTRAT), ("max_dd", C.c_double * NSTRAT)]


class MctsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmcts error %d: %s" % (code, msg))
        self.code = code


# every symbol include/mcts.h declares: (name, restype, argtypes)
_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
PROTOTYPES = [
    ("mcts_abi_version", C.c_int, []),
    ("mcts_generator_version", C.c_int, []),
    ("mcts_last_error", C.c_char_p, []),
    ("mcts_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("mcts_destroy", None, [C.c_void_p]),
    ("mcts_sequence_length", C.c_int64, [C.POINTER(Params)]),
    ("mcts_simulate", C.c_int, [C.c_void_p, C.POINTER(Params), C.c_uint64, C.POINTER(Options), C.POINTER(Agg), _dp, _dp]),
    ("mcts_replay", C.c_int, [C.c_void_p, _dp, C.c_int64, C.c_int64, C.POINTER(Options), C.POINTER(Agg), _dp, _dp]),
    ("mcts_sequences", C.c_int, [C.c_void_p, C.POINTER(Params), C.c_uint64, C.c_int64, C.c_int64, _dp]),
    ("mcts_plan_count", C.c_int64, [C.c_int64, C.c_int64]),
    ("mcts_plan", C.c_int, [C.c_int64, C.c_int64, _i64p, _i64p]),
    ("mcts_simulate_chunks", C.c_int, [C.c_void_p, C.POINTER(Params), C.c_uint64, C.POINTER(Options), C.c_int64, C.c_int64,
                                       C.POINTER(ChunkResult)]),
    ("mcts_combine", C.c_int, [C.c_int64, C.c_int64, C.POINTER(ChunkResult), C.c_int64, C.POINTER(Agg)]),
    ("mcts_auto_chunk", C.c_int64, [C.c_int64, C.c_int]),
    ("mcts_plan_shards", C.c_int, [C.c_int64, C.c_int64, C.c_int, _i64p]),
    ("mcts_group_create", C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    ("mcts_group_destroy", None, [C.c_void_p]),
    ("mcts_group_size", C.c_int, [C.c_void_p]),
    ("mcts_group_device", C.c_int, [C.c_void_p, C.c_int]),
    ("mcts_group_simulate", C.c_int, [C.c_void_p, C.POINTER(Params), C.c_uint64, C.POINTER(Options), C.POINTER(Agg), _dp, _dp]),
    ("mcts_group_get_timing", C.c_int, [C.c_void_p, C.POINTER(Timing)]),
    ("mcts_export_curves", C.c_int, [C.c_void_p, C.POINTER(Params), C.c_uint64, C.c_int64, C.c_int64, C.c_void_p]),
    ("mcts_device_alloc", C.c_void_p, [C.c_void_p, C.c_int64]),
    ("mcts_device_free", None, [C.c_void_p, C.c_void_p]),
    ("mcts_copy_to_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    ("mcts_summary_rows", C.c_int, [C.POINTER(Agg), C.c_int32, _dp]),
    ("mcts_get_timing", C.c_int, [C.c_void_p, C.POINTER(Timing)]),
    ("mcts_get_stream", C.c_void_p, [C.c_void_p]),
    ("mcts_measure_peaks", C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    ("mcts_measure_step_ceiling", C.c_int, [C.c_void_p, _dp]),
    ("mcts_measure_mix_ceiling", C.c_int, [C.c_void_p, _dp, _dp]),
]

_lib = None


def lib():
    """Load libmcts.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MctsError(ENODEVICE, "%s not found: build it with `python -m mc_trade_sizing_b200.build` "
                                        "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in PROTOTYPES:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.mcts_abi_version() != ABI_VERSION:
            raise MctsError(EINVAL, "libmcts ABI %d != binding ABI %d" % (L.mcts_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


def check(rc):
    if rc < 0:
        msg = lib().mcts_last_error()
        raise MctsError(rc, msg.decode("utf-8", "replace") if msg else "")
    return rc
