"""ctypes loader of tests/cpu_harness (the product's host/device headers compiled with g++).
TEST ONLY -- lets CPU-only CI exercise the exact evaluator / generator / planner code the kernels use."""
import ctypes as C
import os
import subprocess

import numpy as np

from mc_trade_sizing_b200 import _ffi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_harness", "harness.cpp")
LIB = os.path.join(HERE, "cpu_harness", "libharness.so")
CSRC = os.path.join(os.path.dirname(HERE), "mc_trade_sizing_b200", "csrc")

_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
        if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
            tmp = "%s.%d.tmp" % (LIB, os.getpid())   # atomic replace: several test processes may race here
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                            "-Wno-unknown-pragmas", "-o", tmp, SRC], check=True, capture_output=True)
            os.replace(tmp, LIB)
        L = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        L.h_eval_values.restype = C.c_int
        L.h_eval_values.argtypes = [dp, C.c_int64, C.c_int64, C.c_int, dp, dp]
        L.h_simulate.restype = C.c_int
        L.h_simulate.argtypes = [C.POINTER(_ffi.Params), C.c_uint64, C.c_int, dp, dp, dp, C.c_int64, C.c_int64]
        L.h_base_counts.restype = C.c_int
        L.h_base_counts.argtypes = [C.POINTER(_ffi.Params), C.c_uint64, C.c_int64, C.POINTER(C.c_uint32), dp,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.h_philox.restype = None
        L.h_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.h_pairwise.restype = C.c_double
        L.h_pairwise.argtypes = [dp, C.c_int64]
        L.h_chunked_sum.restype = C.c_double
        L.h_chunked_sum.argtypes = [dp, C.c_int64, C.c_int64]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def eval_values(values, dedup=False):
    v = np.ascontiguousarray(values, dtype=np.float64)
    P, T = v.shape
    profit, dd = np.zeros((20, P)), np.zeros((20, P))
    rc = lib().h_eval_values(_dp(v), P, T, int(dedup), _dp(profit), _dp(dd))
    assert rc == 0, rc
    return profit, dd


def simulate(params, seed, first_path=0, n=None, dedup=False, want_eval=True, int_path=False):
    P = params.num_simulations * params.num_mc_shuffles
    n = P - first_path if n is None else n
    L = int(_ffi.lib().mcts_sequence_length(C.byref(params)))
    profit = np.zeros((20, n)) if want_eval else None
    dd = np.zeros((20, n)) if want_eval else None
    seqs = np.zeros((n, L))
    rc = lib().h_simulate(C.byref(params), C.c_uint64(seed), int(dedup) | (2 if int_path else 0), _dp(profit), _dp(dd), _dp(seqs), first_path, n)
    if rc != 0:
        raise ValueError("harness rc %d" % rc)
    return {"profit": profit, "dd": dd, "seqs": seqs}


def base_counts(params, seed, sim):
    counts = (C.c_uint32 * 32)()
    vals = np.zeros(32)
    K, L = C.c_int(), C.c_int()
    rc = lib().h_base_counts(C.byref(params), C.c_uint64(seed), sim, counts, _dp(vals), C.byref(K), C.byref(L))
    if rc != 0:
        raise ValueError("harness rc %d" % rc)
    return np.array(list(counts))[:K.value], vals[:K.value], L.value


def philox(key, ctr):
    out = (C.c_uint32 * 4)()
    lib().h_philox(key[0], key[1], ctr[0], ctr[1], ctr[2], ctr[3], out)
    return list(out)


def pairwise(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().h_pairwise(_dp(a), a.size))


def chunked_sum(a, chunk_paths):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().h_chunked_sum(_dp(a), a.size, chunk_paths))
