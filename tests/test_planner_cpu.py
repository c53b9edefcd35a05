"""CPU tests of the host planner and the C-ABI pieces that need no GPU: symbol export, the pairwise
tree plan / combine (np.mean order across chunks and devices), summary-row arithmetic, errors."""
import ctypes as C

import numpy as np
import pytest

import harness
from conftest import assert_bitexact
from mc_trade_sizing_b200 import Engine, _ffi, make_params
from mc_trade_sizing_b200 import trading_models as tm
from oracle import oracle


def test_library_exports_every_symbol_of_the_header():
    import os
    import re
    lib = _ffi.lib()
    hdr = open(os.path.join(os.path.dirname(_ffi.HERE), "include", "mcts.h")).read()
    declared = set(re.findall(r"\b(mcts_[a-z_0-9]+)\s*\(", hdr))
    assert declared == {name for name, _, _ in _ffi.PROTOTYPES}, declared ^ {n for n, _, _ in _ffi.PROTOTYPES}
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mcts_abi_version() == _ffi.ABI_VERSION
    assert lib.mcts_generator_version() == int(re.search(r"#define MCTS_GENERATOR_VERSION (\d+)", hdr).group(1)) == 3
    assert C.sizeof(_ffi.ChunkResult) == 16 + 6 * 20 * 8 and C.sizeof(_ffi.Agg) == 56


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_ffi.MctsError) as e:
        Engine(0)
    assert e.value.code == _ffi.ENODEVICE
    with pytest.raises(_ffi.MctsError):
        tm.run_all_strategies(0.82, 186.0, 219.0, 1000, 3, 2)


def test_pairwise_restatement_matches_numpy():
    rng = np.random.default_rng(5)
    for n in list(range(1, 40)) + [127, 128, 129, 136, 255, 256, 257, 1000, 1001, 4095, 4104, 20000, 100003]:
        a = rng.normal(scale=1e3, size=n) + rng.integers(-3, 3, size=n) * 204.60000000000002
        assert harness.pairwise(a) == float(np.sum(a)), n


def test_chunked_device_reduction_equals_numpy_mean_order():
    """per-chunk node sums (as k_reduce_nodes / k_reduce_top compute them) + host combine == np.sum"""
    rng = np.random.default_rng(6)
    for P in (1, 7, 8, 100, 128, 129, 1000, 4096, 4097, 50000, 262147, 1 << 20):
        a = rng.normal(scale=5e3, size=P) + 103063.2
        for chunk in (128, 1000, 4096, 65536, 0):
            assert harness.chunked_sum(a, chunk) == float(np.sum(a)), (P, chunk)


def test_plan_tiles_the_path_range_with_tree_nodes():
    for P in (1, 127, 128, 129, 1000, 8388640, 10**9):
        for chunk in (128, 4096, 1 << 21):
            first, count = Engine.plan(P, chunk)
            assert first[0] == 0 and (first[1:] == first[:-1] + count[:-1]).all() and first[-1] + count[-1] == P
            assert count.max() <= max(chunk, 128) and (first % 8 == 0).all()
    first, count = Engine.plan(10**9, 0)
    assert len(first) == 256


def _records_from_per_path(profit, dd, first, count):
    n = len(first)
    recs = (_ffi.ChunkResult * n)()
    for c in range(n):
        lo, hi = int(first[c]), int(first[c] + count[c])
        recs[c].first_path, recs[c].n_paths = lo, hi - lo
        for s in range(20):
            recs[c].sum_profit[s] = harness.pairwise(profit[s, lo:hi]) if hi - lo >= 1 else 0.0
            recs[c].sum_dd[s] = harness.pairwise(dd[s, lo:hi])
            recs[c].min_profit[s], recs[c].max_profit[s] = profit[s, lo:hi].min(), profit[s, lo:hi].max()
            recs[c].min_dd[s], recs[c].max_dd[s] = dd[s, lo:hi].min(), dd[s, lo:hi].max()
    return recs


def test_combine_and_summary_rows_match_the_reference_aggregation():
    """mcts_combine + mcts_summary_rows on per-chunk records == np.mean/np.min/np.max/ratios (:509-521)."""
    prm = make_params(0.82, 186.0, 219.0, 100, 150, 20)
    out = harness.simulate(prm, 3)
    P = 3000
    ref = oracle.summarise(out["profit"], out["dd"], 100)
    for chunk in (128, 500, 0):
        first, count = Engine.plan(P, chunk)
        recs = _records_from_per_path(out["profit"], out["dd"], first, count)
        shuffled = (_ffi.ChunkResult * len(first))(*reversed(list(recs)))   # any order
        agg = Engine.combine(P, chunk, shuffled, len(first))
        from mc_trade_sizing_b200.engine import agg_to_rows
        assert_bitexact(agg_to_rows(agg, 100), ref, "chunk %d" % chunk)
    with pytest.raises(ValueError):
        Engine.combine(P, 128, recs, 3)


def test_ranking_matches_reference_rows(kat):
    """summary_to_rows ranks like :524-528 (stable, ties keep id order) and keeps the description strings."""
    for name in sorted(kat["meta"]):
        ids, rows = kat[name + "/rank_ids"], kat[name + "/rows"]
        by_id = np.zeros((20, 9))
        by_id[ids - 1] = rows
        ranked = tm.summary_to_rows(by_id)
        assert [tm.DESCRIPTIONS[int(i)] for i in ids] == [r[0] for r in ranked], name
        assert_bitexact(np.array([r[1:] for r in ranked]), rows, name)
    assert list(kat["descriptions"]) == [tm.DESCRIPTIONS[i] for i in range(1, 21)]


def test_parameter_validation_mirrors_numpy():
    lib = _ffi.lib()
    assert lib.mcts_sequence_length(C.byref(make_params(0.8, 1.0, 1.0, 1, 2, 2, use_markov2=True))) == 2
    assert lib.mcts_sequence_length(C.byref(make_params(0.8, 1.0, 1.0, 0, 2, 2))) == 0
    assert lib.mcts_sequence_length(C.byref(make_params(0.8, 1.0, 1.0, 1200, 2, 2, use_regime=True, regimes=[
        {"length": 300, "hit_rate": 0.9, "avg_win": 200, "avg_loss": 100}]))) == 300
    assert lib.mcts_sequence_length(C.byref(make_params(1.3, 1.0, 1.0, 10, 2, 2))) == _ffi.EINVAL   # p > 1 in choice()
    assert b"ValueError" in lib.mcts_last_error()
    assert lib.mcts_sequence_length(C.byref(make_params(1.3, 1.0, 1.0, 10, 2, 2, use_markov=True))) == 10  # rand() < p never validates


def test_simulate_refuses_sums_beyond_dbl_max():
    """3 * max|value| * trades must stay finite, otherwise inf - inf = NaN could arise in the drawdown, which
    the reference reports (np.min propagates NaN) and the kernels' ordered comparisons would drop."""
    lib = _ffi.lib()
    assert lib.mcts_sequence_length(C.byref(make_params(0.5, 1e304, 1e304, 1000, 2, 2))) == 1000      # 3e307: fine
    assert lib.mcts_sequence_length(C.byref(make_params(0.5, 1e305, 1e305, 1000, 2, 2))) == _ffi.EUNSUPPORTED
    assert b"DBL_MAX" in lib.mcts_last_error()
    assert lib.mcts_sequence_length(C.byref(make_params(0.5, 1e305, 1e305, 10, 2, 2))) == 10


def test_group_api_needs_a_device_too():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mc_trade_sizing_b200 import EngineGroup
    for devices in (None, [0], [0, 1]):
        with pytest.raises(_ffi.MctsError) as e:
            EngineGroup(devices)
        assert e.value.code == _ffi.ENODEVICE
