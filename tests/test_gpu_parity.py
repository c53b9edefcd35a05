"""GPU parity tests (run on the B200 box): every call goes through the C ABI of libmcts.so.

  replay   the reference's own post-shuffle sequences (tests/golden/kat.npz, produced by the live
           reference) -> per-path profit / drawdown of all 20 strategies and the summary rows must be
           BIT-EXACT (fp64 compared as uint64 patterns)
  simulate on-device RNG: (a) the per-path results must equal the oracle's evaluation of the very
           sequences the device generated, bit for bit, (b) aggregates must equal NumPy-ordered
           aggregation of those per-path values, bit for bit, (c) results must not depend on chunking,
           (d) distributions must agree with the reference within clustered confidence intervals
           (tolerance stated in the test)
"""
import ctypes as C

import numpy as np
import pytest

import harness
from conftest import assert_bitexact
from mc_trade_sizing_b200 import Engine, make_params, run_all_strategies
from mc_trade_sizing_b200 import _ffi
from oracle import oracle

pytestmark = pytest.mark.gpu

README_REGIMES = [
    {"length": 300, "hit_rate": 0.9, "avg_win": 200, "avg_loss": 100},
    {"length": 200, "hit_rate": 0.5, "avg_win": 100, "avg_loss": 100},
    {"length": 500, "hit_rate": 0.2, "avg_win": 100, "avg_loss": 200},
]
MODE_KW = {
    "dynamic": {},
    "markov1": dict(use_markov=True, p_win_after_win=0.75, p_win_after_loss=0.60),
    "markov2": dict(use_markov2=True, p_win_ww=0.85, p_win_wl=0.65, p_win_lw=0.55, p_win_ll=0.30),
    "regime": dict(use_regime=True, regimes=README_REGIMES),
}
TIE_GROUPS = [(1, 3), (2, 13), (4, 5, 14, 17), (6, 7), (9, 10, 15), (11, 12)]


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


# ---------------------------------------------------------------------------------------------
# replay: bit-exact against the reference
# ---------------------------------------------------------------------------------------------

def test_replay_golden_sequences_bitexact(eng, kat):
    for name in sorted(kat["meta"]):
        meta = kat["meta"][name]
        seqs = kat[name + "/seqs"]
        for dedup in (False, True):
            out = eng.replay(seqs, num_trades=meta["args"][3], options=eng.options(dedup=dedup))
            assert_bitexact(out["profit"], kat[name + "/profit"], "%s profit dedup=%s" % (name, dedup))
            assert_bitexact(out["dd"], kat[name + "/dd"], "%s drawdown dedup=%s" % (name, dedup))
            ids = kat[name + "/rank_ids"]
            assert_bitexact(out["summary"][ids - 1], kat[name + "/rows"], "%s summary rows" % name)


def test_replay_random_values_vs_oracle(eng):
    rng = np.random.default_rng(3)
    pool = np.array([204.60000000000002, 167.4, -240.9, 186.0, -219.0, 0.0, -0.0, 1e-3, -7.25])
    for T in [1, 2, 5, 7, 8, 9, 16, 23, 64, 127, 128, 129, 136, 200, 255, 256, 257, 264, 1000, 1001, 1007, 1100, 2049, 5000]:
        P = 70 if T < 2000 else 9
        vals = rng.choice(pool, size=(P, T))
        vals[0] = rng.normal(size=T) * 100
        vals[1] = rng.integers(-300, 300, size=T)
        ref = oracle.eval_paths(vals)
        out = eng.replay(vals)
        assert_bitexact(out["profit"], ref["profit"], "T=%d profit" % T)
        assert_bitexact(out["dd"], ref["dd"], "T=%d dd" % T)
        assert_bitexact(out["summary"], ref["summary"], "T=%d summary" % T)


def test_replay_chunked_aggregates_match_numpy_order(eng):
    """np.mean over many paths: device chunks are pairwise-tree nodes, so any chunk size is exact."""
    rng = np.random.default_rng(4)
    vals = rng.choice([204.60000000000002, 167.4, -240.9, 186.0, -219.0], size=(5000, 64))
    ref = oracle.eval_paths(vals)
    for chunk in (128, 1000, 4096, 0):
        out = eng.replay(vals, options=eng.options(chunk_paths=chunk))
        assert_bitexact(out["profit"], ref["profit"], "chunk %d profit" % chunk)
        assert_bitexact(out["summary"], ref["summary"], "chunk %d summary" % chunk)
        for s in range(20):
            assert out["summary"][s, 0] == float(np.mean(out["profit"][s])), (chunk, s)


def test_replay_staged_copy_many_slices_and_pinned_input(eng):
    """96 MB of caller sequences: twelve 8 MB slices over the stager's worker threads (csrc/stager.hpp), three chunks
    through the two device staging areas; then the same array from pinned memory (direct DMA, no staging)."""
    import torch
    rng = np.random.default_rng(11)
    vals = rng.choice([204.60000000000002, 167.4, -240.9, 186.0, -219.0, 3.5], size=(60000, 200))
    ref = oracle.eval_paths(vals)
    out = eng.replay(vals, options=eng.options(chunk_paths=20000))
    assert_bitexact(out["profit"], ref["profit"], "staged profit")
    assert_bitexact(out["dd"], ref["dd"], "staged dd")
    assert_bitexact(out["summary"], ref["summary"], "staged summary")
    pinned = torch.from_numpy(vals).pin_memory()
    out2 = eng.replay(pinned.numpy(), options=eng.options(chunk_paths=20000))
    assert_bitexact(out2["profit"], ref["profit"], "pinned profit")
    assert_bitexact(out2["summary"], ref["summary"], "pinned summary")


def test_replay_int64_regime_arrays(eng, kat):
    """The reference's regime arrays are int64; fp64 on device is exact for them."""
    seqs = kat["a5_regime/seqs"]
    out = eng.replay(seqs.astype(np.int64), num_trades=1000)
    assert_bitexact(out["profit"], kat["a5_regime/profit"])
    assert_bitexact(out["dd"], kat["a5_regime/dd"])


# ---------------------------------------------------------------------------------------------
# on-device RNG
# ---------------------------------------------------------------------------------------------

def _sim_cases():
    base = (0.82, 186.0, 219.0)
    cases = []
    for mode, kw in MODE_KW.items():
        for T, ns, nsh in ((1000, 40, 20), (37, 9, 5), (8, 5, 3), (5, 4, 3), (1, 3, 2), (129, 7, 4), (1203, 5, 3)):
            kw2 = dict(kw)
            if mode == "regime" and T != 1000:
                kw2["regimes"] = None
            cases.append(("%s_T%d" % (mode, T), base + (T, ns, nsh), kw2))
    cases.append(("dps_h0.72", (0.72, 186.0, 219.0, 1000, 30, 20), {}))
    cases.append(("dps_h0.77", (0.7699999999999999, 186.0, 219.0, 1000, 30, 20), {}))
    cases.append(("nonint_markov1", (0.6, 186.3, 219.7, 500, 13, 3), MODE_KW["markov1"]))
    cases.append(("regime_short", base + (1200, 4, 3), MODE_KW["regime"]))
    cases.append(("regime_trunc", base + (450, 4, 3), MODE_KW["regime"]))
    return cases


def test_simulate_per_path_equals_oracle_on_device_sequences(eng):
    for name, args, kw in _sim_cases():
        prm = make_params(*args, **kw)
        for seed in (12345, 2**63 + 11):
            seqs = eng.sequences(prm, seed=seed)
            out = eng.simulate(prm, seed=seed, per_path=True)
            ref = oracle.eval_paths(seqs, num_trades=args[3])
            assert_bitexact(out["profit"], ref["profit"], name + " profit")
            assert_bitexact(out["dd"], ref["dd"], name + " dd")
            assert_bitexact(out["summary"], ref["summary"], name + " summary")
            for opts in (eng.options(dedup=True), eng.options(int_fast_path=False), eng.options(dedup=True, int_fast_path=False)):
                alt = eng.simulate(prm, seed=seed, per_path=True, options=opts)   # int32 / fp64 evaluators, 20 ids / 11 classes
                assert_bitexact(alt["profit"], out["profit"], name + " variant profit")
                assert_bitexact(alt["dd"], out["dd"], name + " variant dd")
                assert_bitexact(alt["summary"], out["summary"], name + " variant summary")


def test_device_generator_matches_its_cpu_model(eng):
    """Sequences from the GPU == the same Philox/urn code run on the host (partition invariance: any
    path range regenerates identically)."""
    for name, args, kw in _sim_cases():
        prm = make_params(*args, **kw)
        dev = eng.sequences(prm, seed=777)
        host = harness.simulate(prm, 777, want_eval=False)["seqs"]
        assert_bitexact(dev, host, name)
        P = args[4] * args[5]
        if P > 5:
            part = eng.sequences(prm, seed=777, first_path=3, n=P - 5)
            assert_bitexact(part, dev[3:P - 2], name + " sub-range")


def test_device_generator_matches_independent_restatement(eng):
    """GPU sequences == oracle/devgen_ref.c (plain C written from the generator specification)."""
    many = [{"length": 40 + 3 * i, "hit_rate": 0.3 + 0.04 * i, "avg_win": 50.0 + 10 * i, "avg_loss": 20.5 + 7 * i} for i in range(7)]
    wide = [{"length": 20 + i, "hit_rate": 0.2 + 0.04 * i, "avg_win": 10.0 + 3 * i, "avg_loss": 5.5 + 2 * i} for i in range(16)]
    cases = _sim_cases() + [("regime_14_values", (0.82, 186.0, 219.0, 400, 30, 5), dict(use_regime=True, regimes=many)),
                            ("regime_32_values_fp64_fallback", (0.82, 186.0, 219.0, 440, 40, 5), dict(use_regime=True, regimes=wide))]
    for name, args, kw in cases:
        dev = eng.sequences(make_params(*args, **kw), seed=31337)
        ref = oracle.devgen_sequences(oracle.make_params(*args, **kw), 31337)
        assert_bitexact(dev, ref, name)
        out = eng.simulate(make_params(*args, **kw), seed=31337, per_path=True)
        exp = oracle.eval_paths(ref, num_trades=args[3])
        assert_bitexact(out["profit"], exp["profit"], name + " profit")
        assert_bitexact(out["dd"], exp["dd"], name + " dd")
        assert_bitexact(out["summary"], exp["summary"], name + " summary")


def test_fuzz_random_parameter_sets(eng):
    """400 random parameter sets (degenerate probabilities, zero / negative / equal / tiny values, random
    regime lists, trade counts 1..1203): GPU sequences == devgen_ref.c, GPU per-path results and
    aggregates == oracle on those sequences (all evaluator variants), and errors for the same inputs."""
    import fuzz_cases
    ok = err = 0
    for args, kw in fuzz_cases.random_cases(400, 777):
        prm = make_params(*args, **kw)
        try:
            ref_seqs = oracle.devgen_sequences(oracle.make_params(*args, **kw), 99)
        except ValueError:
            ref_seqs = None
        if ref_seqs is None or ref_seqs.shape[1] == 0:
            with pytest.raises(ValueError):
                eng.simulate(prm, seed=99)
            err += 1
            continue
        opts = eng.options(dedup=bool(ok & 1), int_fast_path=bool(ok & 2), chunk_paths=128 if ok & 4 else 0)
        try:
            out = eng.simulate(prm, seed=99, per_path=True, options=opts)
        except _ffi.MctsError as e:
            assert e.code == _ffi.EUNSUPPORTED, (args, kw)
            vals = np.unique(ref_seqs)
            assert np.any((vals != 0) & (np.abs(vals) < 4.5e-308)), (args, kw)   # only subnormal values are refused
            continue
        assert_bitexact(eng.sequences(prm, seed=99), ref_seqs, "%s %s" % (args, kw))
        exp = oracle.eval_paths(ref_seqs, num_trades=args[3])
        assert_bitexact(out["profit"], exp["profit"], "%s %s profit" % (args, kw))
        assert_bitexact(out["dd"], exp["dd"], "%s %s dd" % (args, kw))
        if args[3] > 0:
            assert_bitexact(out["summary"], exp["summary"], "%s %s summary" % (args, kw))
        ok += 1
    assert ok > 250 and err >= 5, (ok, err)


def test_long_paths_use_the_global_code_store(eng):
    """Paths too long for shared-memory code columns (> ~3,500 trades) keep them in global memory: same
    generator specification, same evaluator, same results."""
    for T, ns, nsh, kw in ((3000, 6, 4, {}), (5000, 5, 3, {}), (20001, 3, 2, MODE_KW["markov1"]), (40000, 2, 2, {}),
                           (7000, 4, 3, dict(use_regime=True, regimes=None))):
        args = (0.82, 186.0, 219.0, T, ns, nsh)
        prm = make_params(*args, **kw)
        ref_seqs = oracle.devgen_sequences(oracle.make_params(*args, **kw), 2718)
        assert_bitexact(eng.sequences(prm, seed=2718), ref_seqs, "T=%d sequences" % T)
        exp = oracle.eval_paths(ref_seqs, num_trades=T)
        for opts in (eng.options(), eng.options(dedup=True, int_fast_path=False)):
            out = eng.simulate(prm, seed=2718, per_path=True, options=opts)
            assert_bitexact(out["profit"], exp["profit"], "T=%d profit" % T)
            assert_bitexact(out["dd"], exp["dd"], "T=%d dd" % T)
            assert_bitexact(out["summary"], exp["summary"], "T=%d summary" % T)


def test_sequences_are_permutations_of_one_multiset_per_simulation(eng):
    prm = make_params(0.82, 186.0, 219.0, 1000, 6, 5)
    seqs = eng.sequences(prm, seed=5)
    for sim in range(6):
        block = np.sort(seqs[sim * 5:(sim + 1) * 5], axis=1)
        assert (block == block[0]).all()
        assert not (seqs[sim * 5] == seqs[sim * 5 + 1]).all()
    assert not (np.sort(seqs[0]) == np.sort(seqs[5])).all()


def test_simulate_independent_of_chunking(eng):
    prm = make_params(0.82, 186.0, 219.0, 200, 700, 7)
    ref = eng.simulate(prm, seed=9, per_path=True)
    for chunk in (128, 777, 2048):
        out = eng.simulate(prm, seed=9, per_path=True, options=eng.options(chunk_paths=chunk))
        assert_bitexact(out["profit"], ref["profit"], "chunk %d" % chunk)
        assert_bitexact(out["dd"], ref["dd"], "chunk %d" % chunk)
        assert_bitexact(out["summary"], ref["summary"], "chunk %d" % chunk)
    assert_bitexact(ref["summary"], oracle.summarise(ref["profit"], ref["dd"], 200), "np.mean order")


def test_multi_device_split_combines_exactly(eng):
    """Emulate G ranks on one GPU: disjoint chunk ranges, host-side tree combine."""
    prm = make_params(0.82, 186.0, 219.0, 100, 900, 10)
    P, chunk = 9000, 512
    ref = eng.simulate(prm, seed=21, options=eng.options(chunk_paths=chunk))
    first, count = Engine.plan(P, chunk)
    n = len(first)
    for G in (2, 3, 8):
        recs = (_ffi.ChunkResult * n)()
        bounds = [n * g // G for g in range(G + 1)]
        for g in reversed(range(G)):  # any order
            part, m = eng.simulate_chunks(prm, 21, bounds[g], bounds[g + 1], options=eng.options(chunk_paths=chunk))
            for i in range(m):
                C.memmove(C.byref(recs[bounds[g] + i]), C.byref(part[i]), C.sizeof(_ffi.ChunkResult))
        agg = Engine.combine(P, chunk, recs, n)
        for s in range(20):
            for f, _ in _ffi.Agg._fields_:
                a, b = getattr(agg[s], f), getattr(ref["agg"][s], f)
                assert np.float64(a).view(np.uint64) == np.float64(b).view(np.uint64), (G, s, f)


def _cluster_se(x, n_sh):
    """standard error of the mean with paths clustered by simulation (shared multiset)."""
    m = x.reshape(-1, n_sh).mean(axis=1)
    return m.std(ddof=1) / np.sqrt(m.size)


def test_distribution_agrees_with_reference(eng, stats_golden):
    """Per-strategy mean profit / mean drawdown from the device RNG vs the live reference's sample
    (tests/golden/stats.npz: 300 sims x 4 shuffles per mode).  Tolerance: |difference| <= 4.5 combined
    clustered standard errors (two-sided p ~ 7e-6 per comparison, 160 comparisons)."""
    for mode, kw in MODE_KW.items():
        meta = stats_golden["meta"][mode]
        n_sh = meta["args"][5]
        prm = make_params(0.82, 186.0, 219.0, 1000, 20000, n_sh, **kw)
        out = eng.simulate(prm, seed=4242, per_path=True)
        for s in range(20):
            for key in ("profit", "dd"):
                ref, dev = stats_golden["%s/%s" % (mode, key)][s], out[key][s]
                se = np.hypot(_cluster_se(ref, n_sh), _cluster_se(dev, n_sh))
                assert abs(ref.mean() - dev.mean()) <= 4.5 * se + 1e-9, (mode, s + 1, key, ref.mean(), dev.mean(), se)
        # spread of the per-path distribution: quantiles within 4.5 bootstrap SE of the reference sample
        rng = np.random.default_rng(0)
        for s in (0, 1, 18):
            ref, dev = stats_golden["%s/profit" % mode][s], out["profit"][s]
            for q in (0.1, 0.5, 0.9):
                boots = [np.quantile(rng.choice(ref.reshape(-1, n_sh), size=ref.size // n_sh).ravel(), q) for _ in range(200)]
                assert abs(np.quantile(ref, q) - np.quantile(dev, q)) <= 4.5 * np.std(boots) + 1e-9, (mode, s + 1, q)


def test_distribution_agrees_with_seeded_reference_port(eng):
    """Larger two-sample check: the oracle (bit-exact restatement of the reference incl. MT19937, so its
    samples ARE reference samples) vs the device RNG, 8000 simulations x 4 shuffles per mode.
    Tolerances: means within 4.5 clustered standard errors; two-sample Kolmogorov-Smirnov distance of the
    per-simulation means below the 1e-5 critical value c(alpha)*sqrt(2/n), c = 1.95 * 1.25 safety."""
    n_sims, n_sh = 8000, 4
    for mode, kw in MODE_KW.items():
        args = (0.82, 186.0, 219.0, 1000, n_sims, n_sh)
        ref = oracle.run_all(oracle.make_params(*args, **kw), 987654, per_path=True)
        dev = eng.simulate(make_params(*args, **kw), seed=13, per_path=True)
        for s in range(20):
            for key in ("profit", "dd"):
                a, b = ref[key][s], dev[key][s]
                se = np.hypot(_cluster_se(a, n_sh), _cluster_se(b, n_sh))
                assert abs(a.mean() - b.mean()) <= 4.5 * se + 1e-9, (mode, s + 1, key, a.mean(), b.mean(), se)
                ca, cb = np.sort(a.reshape(-1, n_sh).mean(axis=1)), np.sort(b.reshape(-1, n_sh).mean(axis=1))
                grid = np.concatenate([ca, cb])
                ks = np.max(np.abs(np.searchsorted(ca, grid, side="right") / ca.size - np.searchsorted(cb, grid, side="right") / cb.size))
                assert ks <= 1.95 * 1.25 * np.sqrt(2.0 / n_sims), (mode, s + 1, key, ks)


def test_analytic_anchor_strategy1(eng):
    """expected profit of strategy 1 from the chains' exact win-count expectations."""
    anchors = {"dynamic": 103063.2, "markov1": 66936.7, "markov2": 57802.5, "regime": -19000.0}
    for mode, kw in MODE_KW.items():
        prm = make_params(0.82, 186.0, 219.0, 1000, 100000, 1, **kw)
        out = eng.simulate(prm, seed=1, per_path=True)
        p = out["profit"][0]
        se = p.std(ddof=1) / np.sqrt(p.size)
        assert abs(p.mean() - anchors[mode]) <= 4.5 * se + 0.06, (mode, p.mean(), anchors[mode], se)


def test_full_size_markov2_properties(eng):
    """Config C3 at full size (1e6 paths): size-independent properties."""
    prm = make_params(0.82, 186.0, 219.0, 1000, 50000, 20, **MODE_KW["markov2"])
    out = eng.simulate(prm, seed=99, per_path=True)
    p, d = out["profit"], out["dd"]
    for g in TIE_GROUPS:
        for other in g[1:]:
            assert_bitexact(p[g[0] - 1], p[other - 1], "tie group %s" % (g,))
            assert_bitexact(d[g[0] - 1], d[other - 1], "tie group %s" % (g,))
    assert (p[0].reshape(-1, 20) == p[0].reshape(-1, 20)[:, :1]).all()  # strategy 1 is shuffle-invariant (integers)
    assert (d <= 0).all() and (p == np.round(p)).all()
    assert (np.abs(p[1]) <= 2 * 219 * 1000).all() and (np.abs(p[17]) <= 3 * 219 * 1000).all()
    assert_bitexact(out["summary"], oracle.summarise(p, d, 1000), "aggregates in np.mean order")
    t = eng.timing()
    assert t["trade_steps"] == 1000000 * 1000 * 20 and t["launches"] > 0


def test_full_size_regime_config_c4(eng):
    """Config C4 at full size: 1e7 regime paths (README regimes).  Aggregates only (per-path arrays would
    be 3.2 GB on the host); size-independent properties: independence of chunking and of the evaluator
    (int32 vs fp64, 20 ids vs 11 classes), tie groups, bounds, analytic mean."""
    prm = make_params(0.82, 186.0, 219.0, 1000, 500000, 20, **MODE_KW["regime"])
    ref = eng.simulate(prm, seed=2026)
    for opts in (eng.options(chunk_paths=1 << 20), eng.options(dedup=True), eng.options(int_fast_path=False, chunk_paths=1 << 19)):
        assert_bitexact(eng.simulate(prm, seed=2026, options=opts)["summary"], ref["summary"], "C4 variant")
    s = ref["summary"]
    for g in TIE_GROUPS:
        for other in g[1:]:
            assert_bitexact(s[g[0] - 1], s[other - 1], "tie group %s" % (g,))
    assert (s[:, 5] <= s[:, 6]).all() and (s[:, 6] <= 0).all() and (s[:, 3] <= s[:, 0]).all() and (s[:, 0] <= s[:, 4]).all()
    assert abs(s[0, 0] - (-19000.0)) < 5 * 3400 / np.sqrt(500000)      # SURVEY A.1 anchor, sd per sim ~3400
    assert ref["agg"][0].count == 10**7


def test_scale_sweep_config_c5_slice(eng):
    """Config C5 (1e9 'No Markov' paths over 1/2/4/8 GPUs) on a 2e7-path slice: non-integer values, so the mean depends on
    the summation order.  Size-independent properties: the aggregates are bit-identical for the library-default chunking,
    for 2 / 3 / 8 emulated ranks (disjoint chunk ranges combined on the host) and for the dedup evaluator; tie groups;
    the analytic mean of strategy 1."""
    n_sims, n_sh = 1_000_000, 20
    prm = make_params(0.82, 186.0, 219.0, 1000, n_sims, n_sh)
    P = n_sims * n_sh
    ref = eng.simulate(prm, seed=5)
    s = ref["summary"]
    assert_bitexact(eng.simulate(prm, seed=5, options=eng.options(dedup=True))["summary"], s, "dedup")
    first, count = Engine.plan(P, 0)
    n = len(first)
    assert n >= 8
    for G in (2, 3, 8):
        recs = (_ffi.ChunkResult * n)()
        bounds = [n * g // G for g in range(G + 1)]
        for g in range(G):
            part, m = eng.simulate_chunks(prm, 5, bounds[g], bounds[g + 1])
            for i in range(m):
                C.memmove(C.byref(recs[bounds[g] + i]), C.byref(part[i]), C.sizeof(_ffi.ChunkResult))
        agg = Engine.combine(P, 0, recs, n)
        for k in range(20):
            for f, _ in _ffi.Agg._fields_:
                a, b = getattr(agg[k], f), getattr(ref["agg"][k], f)
                assert np.float64(a).view(np.uint64) == np.float64(b).view(np.uint64), (G, k, f)
    for g in TIE_GROUPS:
        for other in g[1:]:
            assert_bitexact(s[g[0] - 1], s[other - 1], "tie group %s" % (g,))
    assert abs(s[0, 0] - 103063.2) < 5 * 6300 / np.sqrt(n_sims)      # sd of a simulation's strategy-1 profit ~6300
    assert ref["agg"][0].count == P


def test_curve_export_matches_reference_equity_curves(eng):
    """Extension K4: exported curves == np.cumsum of each strategy's contributions on the device-generated
    sequences (the temporary of :167), bit for bit; layout [20][ceil(L/2)][n][2]."""
    for name, args, kw in [("dynamic", (0.82, 186.0, 219.0, 1000, 10, 20), {}), ("odd_len", (0.82, 186.0, 219.0, 37, 9, 5), {}),
                           ("regime", (0.82, 186.0, 219.0, 1000, 6, 4), MODE_KW["regime"]), ("markov2", (0.82, 186.0, 219.0, 129, 7, 4), MODE_KW["markov2"])]:
        prm = make_params(*args, **kw)
        P = args[4] * args[5]
        first, n = 3, P - 5
        seqs = eng.sequences(prm, seed=8, first_path=first, n=n)
        L = seqs.shape[1]
        cur = eng.export_curves(prm, 8, first, n)                      # [20, half, n, 2]
        got = cur.transpose(0, 2, 1, 3).reshape(20, n, -1)[:, :, :L]   # [20, n, L]
        assert_bitexact(got, oracle.equity_curves(seqs), name)
        if L % 2:
            assert (cur[:, -1, :, 1] == 0).all()


def test_curve_parquet_round_trip(eng, tmp_path):
    import pyarrow.parquet as pq
    from mc_trade_sizing_b200 import curves
    prm = make_params(0.82, 186.0, 219.0, 37, 4, 3)
    fn = str(tmp_path / "curves.parquet")
    n_rows = curves.write_parquet(eng, prm, 5, 2, 7, fn, strategies=[1, 19])
    tbl = pq.read_table(fn).to_pandas()
    assert n_rows == len(tbl) == 2 * 7 * 37
    ref = oracle.equity_curves(eng.sequences(prm, seed=5, first_path=2, n=7))
    got = tbl[tbl.strategy_id == 19].sort_values(["path", "trade"]).equity.to_numpy().reshape(7, 37)
    assert_bitexact(got, ref[18], "parquet curves of strategy 19")
    assert set(tbl.path) == set(range(2, 9))


# ---------------------------------------------------------------------------------------------
# boundary behaviour
# ---------------------------------------------------------------------------------------------

def test_run_all_strategies_drop_in(eng):
    rows = run_all_strategies(0.82, 186.0, 219.0, 1000, 10, 5, seed=3)
    assert len(rows) == 20 and all(len(r) == 10 and isinstance(r[0], str) for r in rows)
    keys = [r[9] for r in rows]
    pos = [k for k in keys if k >= 0]
    assert keys[:len(pos)] == sorted(pos, reverse=True)
    again = run_all_strategies(0.82, 186.0, 219.0, 1000, 10, 5, seed=3)
    assert rows == again
    prm = make_params(0.82, 186.0, 219.0, 1000, 10, 5)
    per = eng.simulate(prm, seed=3, per_path=True)
    ref = oracle.summarise(per["profit"], per["dd"], 1000)
    by_desc = {r[0]: r[1:] for r in rows}
    from mc_trade_sizing_b200.trading_models import DESCRIPTIONS
    for i in range(20):
        assert_bitexact(np.array(by_desc[DESCRIPTIONS[i + 1]]), ref[i], "row %d" % (i + 1))


def test_three_concurrent_cli_processes_share_the_gpu():
    """dps.py runs three CLI subprocesses at a time (dps.py:268): each creates its own handle and CUDA context on the
    same device; all must succeed and, with the same seed, print the same table."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "src", "trading_models.py"), "--hit_rate", "0.82", "--avg_win", "186", "--avg_loss", "219",
           "--num_simulations", "200", "--num_trades", "1000", "--num_mc_shuffles", "20"]
    env = dict(os.environ, MCTS_SEED="77")
    procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for _ in range(3)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (out, err) in zip(procs, outs):
        assert p.returncode == 0, err[-2000:]
        assert out.count("\n") > 25
    assert outs[0][0] == outs[1][0] == outs[2][0]


def test_dps_sweep_in_one_process(eng):
    """N2: the 12 runs of dps.py in one process; every block parses with dps.py's row regex (dps.py:356)."""
    import re
    from mc_trade_sizing_b200 import sweep
    cfg = {"hit_rate": 0.82, "avg_win": 186, "avg_loss": 219, "num_simulations": 200, "num_trades": 1000, "num_mc_shuffles": 20,
           "p_win_after_win": 0.8, "p_win_after_loss": 0.4, "p_win_ww": 0.8, "p_win_wl": 0.6, "p_win_lw": 0.5, "p_win_ll": 0.3,
           "regimes": README_REGIMES}
    row = re.compile(r"(.+?)\s+(-?\d+\.\d+)" + r"\s+(-?\d+\.\d+)" * 8)
    runs = sweep.run_dps_sweep(cfg)
    assert len(runs) == 12
    for label, hit_rate, argv, text in runs:
        rows = [m for m in (row.match(ln) for ln in text.split("\n") if len(ln.split()) >= 10) if m]
        assert len(rows) == 20, label
        assert {m.group(1).strip() for m in rows} == set(tm_descriptions())


def tm_descriptions():
    from mc_trade_sizing_b200.trading_models import DESCRIPTIONS
    return DESCRIPTIONS.values()


def test_error_behaviour(eng):
    with pytest.raises(ValueError):  # np.min of an empty list raises in the reference
        eng.simulate(make_params(0.8, 1.0, 1.0, 10, 0, 5))
    with pytest.raises(ValueError):  # empty sequence
        eng.simulate(make_params(0.8, 1.0, 1.0, 0, 2, 2))
    with pytest.raises(ValueError):  # np.random.choice: negative probability
        eng.simulate(make_params(1.3, 1.0, 1.0, 10, 2, 2))
    with pytest.raises(ValueError):
        eng.replay(np.zeros((0, 5)))
    with pytest.raises(_ffi.MctsError):
        Engine(99)
