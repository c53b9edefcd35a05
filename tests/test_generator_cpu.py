"""CPU unit tests of the DEVICE generator code (csrc/gen_core.cuh via tests/cpu_harness): Philox
known answers, base-sequence laws, urn shuffle, counter keying."""
import numpy as np
import pytest

import harness
from conftest import assert_bitexact
from mc_trade_sizing_b200 import make_params
from oracle import oracle

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


def test_philox4x32_10_known_answers():
    """Random123 known-answer vectors (kat_vectors: philox4x32 10)."""
    assert harness.philox((0, 0), (0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert harness.philox((0xffffffff, 0xffffffff), (0xffffffff,) * 4) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert harness.philox((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sequence_lengths_match_reference_generators(kat):
    """L is what the reference generators return (Markov-2 >= 2 trades, short / truncated regimes...)."""
    for name, meta in kat["meta"].items():
        prm = make_params(*meta["args"], **meta["kwargs"])
        _, _, L = harness.base_counts(prm, 1, 0)
        assert L == kat[name + "/seqs"].shape[1], name


def test_base_counts_follow_the_generator_laws():
    """Mean win counts per category over many simulations vs the chains' exact expectations."""
    n = 4000
    prm = make_params(0.82, 186.0, 219.0, 1000, n, 1)
    acc = np.zeros(32)
    for sim in range(n):
        c, vals, L = harness.base_counts(prm, 99, sim)
        assert c.sum() == 1000 and L == 1000
        acc[:len(c)] += c
    by_val = dict(zip(vals, acc[:len(vals)] / n))
    assert by_val[204.60000000000002] == 200 and -197.1 not in by_val  # p = min(0.82+0.2, 1) = 1: the loss value cannot occur, no category
    assert abs(by_val[167.4] - 200 * 0.52) < 4 * np.sqrt(200 * 0.52 * 0.48 / n)
    assert abs(by_val[186.0] - 600 * 0.82) < 4 * np.sqrt(600 * 0.82 * 0.18 / n)
    # Markov-1 stationary-ish win fraction: p = 0.6 / (1 - 0.75 + 0.6)
    prm = make_params(0.82, 186.0, 219.0, 1000, n, 1, **MODE_KW["markov1"])
    wins = np.array([harness.base_counts(prm, 5, s)[0][0] for s in range(1500)])
    assert abs(wins.mean() / 1000 - 0.6 / 0.85) < 0.003


def test_paths_are_uniform_permutations_of_the_simulation_multiset():
    prm = make_params(0.82, 186.0, 219.0, 64, 3, 400)
    seqs = harness.simulate(prm, 7, want_eval=False)["seqs"]
    for sim in range(3):
        block = seqs[sim * 400:(sim + 1) * 400]
        srt = np.sort(block, axis=1)
        assert (srt == srt[0]).all()
        # every position sees every value with the multiset's frequency (uniformity, 5 sigma)
        for v in np.unique(block[0]):
            frac = (block[0] == v).mean()
            per_pos = (block == v).mean(axis=0)
            assert np.all(np.abs(per_pos - frac) < 5 * np.sqrt(frac * (1 - frac) / 400) + 1e-12)


def test_counter_keying_is_partition_invariant():
    prm = make_params(0.82, 186.0, 219.0, 203, 9, 7, **MODE_KW["regime"] | {"regimes": None})
    full = harness.simulate(prm, 123456789012345)
    part = harness.simulate(prm, 123456789012345, first_path=17, n=30)
    assert_bitexact(part["seqs"], full["seqs"][17:47])
    assert_bitexact(part["profit"], full["profit"][:, 17:47])
    other = harness.simulate(prm, 123456789012346, first_path=17, n=30)
    assert not (other["seqs"] == part["seqs"]).all()


def test_simulated_paths_evaluate_like_the_reference():
    """device-generated sequences + device evaluator (categorical code path) == oracle on the same sequences"""
    for mode, kw in MODE_KW.items():
        for T in (1000, 129, 37, 8, 5, 1):
            kw2 = dict(kw)
            if mode == "regime" and T != 1000:
                kw2["regimes"] = None
            prm = make_params(0.82, 186.0, 219.0, T, 5, 4, **kw2)
            out = harness.simulate(prm, 12345)
            ref = oracle.eval_paths(out["seqs"])
            assert_bitexact(out["profit"], ref["profit"], "%s T=%d profit" % (mode, T))
            assert_bitexact(out["dd"], ref["dd"], "%s T=%d dd" % (mode, T))


def test_device_generator_code_matches_independent_restatement():
    """csrc/gen_core.cuh (compiled for the host) vs oracle/devgen_ref.c, written from the spec: base
    multiset laws, category order, Fisher-Yates / urn shuffle, counter layout -- bit for bit."""
    assert oracle.devgen_philox((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    many = [{"length": 40 + 3 * i, "hit_rate": 0.3 + 0.04 * i, "avg_win": 50.0 + 10 * i, "avg_loss": 20.5 + 7 * i} for i in range(7)]
    cases = []
    for mode, kw in MODE_KW.items():
        for T in (1000, 129, 37, 9, 8, 5, 2, 1):
            kw2 = dict(kw)
            if mode == "regime" and T != 1000:
                kw2["regimes"] = None
            cases.append(((0.82, 186.0, 219.0, T, 4, 3), kw2))
    cases.append(((0.6, 186.3, 219.7, 500, 3, 3), MODE_KW["markov1"]))
    cases.append(((0.82, 186.0, 219.0, 1200, 2, 3), MODE_KW["regime"]))
    cases.append(((0.82, 186.0, 219.0, 450, 2, 3), MODE_KW["regime"]))
    cases.append(((0.82, 186.0, 219.0, 400, 3, 2), dict(use_regime=True, regimes=many)))      # 14 distinct values
    wide = [{"length": 20 + i, "hit_rate": 0.2 + 0.04 * i, "avg_win": 10.0 + 3 * i, "avg_loss": 5.5 + 2 * i} for i in range(13)]
    cases.append(((0.82, 186.0, 219.0, 330, 3, 3), dict(use_regime=True, regimes=wide)))      # 26 distinct values: fp64 fallback
    cases.append(((0.2, 50.0, 75.5, 333, 3, 4), {}))
    cases.append(((0.5, 10.0, -10.0, 50, 3, 2), dict(use_markov=True)))                           # one distinct value
    for args, kw in cases:
        for seed in (12345, 2**63 + 11):
            dev = harness.simulate(make_params(*args, **kw), seed, want_eval=False)["seqs"]
            ref = oracle.devgen_sequences(oracle.make_params(*args, **kw), seed)
            assert_bitexact(dev, ref, "%s %s" % (args, list(kw)))
    prm = oracle.make_params(0.82, 186.0, 219.0, 203, 9, 7)
    assert_bitexact(oracle.devgen_sequences(prm, 5, first_path=17, n=30), oracle.devgen_sequences(prm, 5)[17:47])


def test_int32_evaluator_is_exact_for_integer_runs():
    """All-integer runs (Markov defaults, README regimes): the int32 evaluator must give the fp64 results."""
    for mode in ("markov1", "markov2", "regime"):
        for T in (1000, 129, 37, 8, 5, 1):
            kw = dict(MODE_KW[mode])
            if mode == "regime" and T != 1000:
                kw["regimes"] = None
            prm = make_params(0.82, 186.0, 219.0, T, 5, 4, **kw)
            ref = oracle.eval_paths(harness.simulate(prm, 99, want_eval=False)["seqs"])
            for dedup in (False, True):
                out = harness.simulate(prm, 99, dedup=dedup, int_path=True)
                assert_bitexact(out["profit"], ref["profit"], "%s T=%d int profit" % (mode, T))
                assert_bitexact(out["dd"], ref["dd"], "%s T=%d int dd" % (mode, T))
    # non-integer values silently stay on the fp64 path
    prm = make_params(0.82, 186.0, 219.0, 100, 3, 2)
    a, b = harness.simulate(prm, 1, int_path=True), harness.simulate(prm, 1)
    assert_bitexact(a["profit"], b["profit"])


def test_fuzz_generator_and_evaluator_against_the_oracles():
    """Random parameter sets (degenerate probabilities, zero / negative / equal / tiny values, random regime
    lists, all trade counts): device code (host-compiled) == devgen_ref.c sequences == oracle evaluation;
    errors are raised for exactly the same inputs."""
    import fuzz_cases
    ok = err = 0
    for args, kw in fuzz_cases.random_cases(400, 2026):
        try:
            ref_seqs = oracle.devgen_sequences(oracle.make_params(*args, **kw), 4242)
            ref_err = None
        except ValueError as e:
            ref_err = e
        prm = make_params(*args, **kw)
        if ref_err is not None or ref_seqs.shape[1] == 0:
            try:
                harness.simulate(prm, 4242)
                raised = False
            except ValueError:
                raised = True
            assert raised, (args, kw)
            err += 1
            continue
        try:
            out = harness.simulate(prm, 4242, int_path=bool(ok & 1), dedup=bool(ok & 2))
        except ValueError:
            # the only inputs the reference accepts but this build rejects: values outside [2^-1021, 2^1023]
            vals = np.unique(ref_seqs)
            assert np.any((vals != 0) & (np.abs(vals) < 4.5e-308)), (args, kw)   # only subnormal values are refused
            continue
        assert_bitexact(out["seqs"], ref_seqs, "%s %s" % (args, kw))
        exp = oracle.eval_paths(ref_seqs)
        assert_bitexact(out["profit"], exp["profit"], "%s %s profit" % (args, kw))
        assert_bitexact(out["dd"], exp["dd"], "%s %s dd" % (args, kw))
        ok += 1
    assert ok > 250 and err >= 5, (ok, err)


def _draw53(seed, sim, i):
    """The i-th 53-bit uniform of a simulation's base stream by the published rule, and its 16-bit first look."""
    key = (seed & 0xffffffff, seed >> 32)
    w = harness.philox(key, (i >> 3, 0, sim, 0))[(i & 7) >> 1]
    h = (w & 0xffff) if i & 1 else w >> 16
    y = harness.philox(key, (i, 1, sim, 0))
    return (h << 37) | (((y[0] << 32) | y[1]) >> 27), h


@pytest.mark.parametrize("draw", [0, 1, 5, 8, 11, 22, 23])
def test_base_draw_tie_of_the_high_bits_uses_the_second_block(draw):
    """A Bernoulli draw of the base stream is decided by the 16-bit first look at its 53-bit uniform unless that ties with the
    threshold's top 16 bits (probability 2^-16); only then is the second Philox block generated.  Force the tie at draw `draw`:
    take the threshold from the very uniform that draw will see -- equal (no win), one above (win), far above / below with the
    first look still tied -- and compare the win counts of all three generator loops (segments: whole blocks of eight; Markov-1
    and Markov-2: per-draw thresholds) with the independent restatement."""
    seed, sim, L = 987654321, 17, 24
    assert harness.philox((1, 2), (3, 4, 5, 6)) == oracle.devgen_philox((1, 2), (3, 4, 5, 6))
    k, h = _draw53(seed, sim, draw)
    tried = 0
    for thr in (k, k + 1, k + 2 ** 30, k - 2 ** 30, (h << 37), (h << 37) + 1):
        if thr >> 37 != h or thr <= 0:
            continue                                      # keep the first look tied
        tried += 1
        p = thr / 2.0 ** 53                               # threshold53(p) == thr exactly (thr < 2^53 is representable)
        for kw in (dict(use_regime=True, regimes=[_regime(L, p, 186.0, 219.0)]),
                   dict(use_markov=True, p_win_after_win=p, p_win_after_loss=p),
                   dict(use_markov2=True, p_win_ww=p, p_win_wl=p, p_win_lw=p, p_win_ll=p)):
            counts, vals, n = harness.base_counts(make_params(p, 186.0, 219.0, L, 1, 1, **kw), seed, sim)
            ref = oracle.devgen_sequences(oracle.make_params(p, 186.0, 219.0, L, sim + 1, 1, **kw), seed, first_path=sim, n=1)[0]
            wins = int(sum(c for c, v in zip(counts, vals) if v > 0))
            assert n == L and wins == int((ref > 0).sum()), (draw, thr, kw, counts)
            expected = sum(1 for i in range(L) if _draw53(seed, sim, i)[0] < thr)   # every draw sees the same threshold
            assert wins == expected, (draw, thr, kw)
    assert tried >= 3


def _regime(length, hit, win, loss):
    return {"length": length, "hit_rate": hit, "avg_win": win, "avg_loss": loss}


URN_EDGE_CASES = [
    ("K1_all_wins", (1.0, 10.0, 7.0, 37, 4, 3), dict(use_regime=True, regimes=[_regime(37, 1.0, 10.0, 7.0)])),               # one value: 2-urn, empty second category
    ("K2_regime", (0.5, 1.0, 1.0, 50, 4, 3), dict(use_regime=True, regimes=[_regime(50, 0.4, 12.5, 3.25)])),
    ("K3_p1_segment", (0.5, 1.0, 1.0, 64, 4, 3), dict(use_regime=True, regimes=[_regime(30, 1.0, 5.0, 9.0), _regime(34, 0.5, 6.5, 2.0)])),    # 5-urn padded
    ("K4_two_regimes", (0.5, 1.0, 1.0, 70, 4, 3), dict(use_regime=True, regimes=[_regime(30, 0.7, 5.0, 9.0), _regime(40, 0.5, 6.5, 2.0)])),  # 5-urn padded
    ("K5_default_dynamic", (0.82, 186.0, 219.0, 100, 4, 3), {}),                                                              # phase 1 has p = 1: five values
    ("K6_dynamic", (0.6, 186.0, 219.0, 100, 4, 3), {}),                                                                       # 6-urn
    ("K6_three_regimes", (0.5, 1.0, 1.0, 90, 4, 3), dict(use_regime=True, regimes=[_regime(30, 0.7, 5.0, 9.0), _regime(30, 0.5, 6.5, 2.0), _regime(30, 0.2, 1.5, 2.5)])),
    ("K7_fisher_yates", (0.5, 1.0, 1.0, 120, 4, 3), dict(use_regime=True, regimes=[_regime(30, 0.7, 5.0, 9.0), _regime(30, 0.5, 6.5, 2.0), _regime(30, 0.2, 1.5, 2.5),
                                                                                        _regime(30, 0.0, 8.0, 4.0)])),         # the 4th regime never wins: 7 values
    ("K8_fisher_yates", (0.5, 1.0, 1.0, 121, 4, 3), dict(use_regime=True, regimes=[_regime(30, 0.7, 5.0, 9.0), _regime(30, 0.5, 6.5, 2.0), _regime(30, 0.2, 1.5, 2.5),
                                                                                        _regime(31, 0.3, 8.0, 4.0)])),
]


@pytest.mark.parametrize("name,args,kw", URN_EDGE_CASES, ids=[c[0] for c in URN_EDGE_CASES])
def test_generator_choice_at_the_value_count_boundaries(name, args, kw):
    """Urn (<= 6 values that can occur, kernels instantiated for 2 / 5 / 6 with empty categories appended) vs Fisher-Yates (more), and
    the rule that impossible outcomes (p >= 1: no loss, p <= 0: no win) get no category: the device code (g++ build) and the
    independent restatement must produce the same sequences, each a permutation of its simulation's multiset."""
    prm = make_params(*args, **kw)
    out = harness.simulate(prm, 321, want_eval=True)
    ref = oracle.devgen_sequences(oracle.make_params(*args, **kw), 321)
    assert_bitexact(out["seqs"], ref, name + " sequences")
    exp = oracle.eval_paths(ref, num_trades=args[3])
    assert_bitexact(out["profit"], exp["profit"], name + " profit")
    n_sh = args[5]
    for sim in range(args[4]):
        rows = np.sort(out["seqs"][sim * n_sh:(sim + 1) * n_sh], axis=1)
        assert (rows == rows[0]).all(), name


def devgen_golden():
    """tests/golden/devgen_v3.npz as {name: (args, kw, sequences)} (made by tests/golden/make_devgen_golden.py from the plain-C restatement)."""
    import os
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_devgen_golden", os.path.join(sys_path, "make_devgen_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    z = np.load(os.path.join(sys_path, "devgen_v3.npz"))
    return mod.SEED, {name: (args, kw, z[name + "__values"][z[name + "__codes"]]) for name, args, kw in mod.CASES}


def test_generator_stream_known_answers():
    """The published stream (MCTS_GENERATOR_VERSION 3) is pinned by committed known answers: the device code (g++ build) and the
    restatement both reproduce every stored sequence.  If this fails after an intended change of the generator, bump the version
    in include/mcts.h and regenerate the fixture."""
    seed, cases = devgen_golden()
    assert len(cases) >= 12
    for name, (args, kw, gold) in cases.items():
        assert_bitexact(oracle.devgen_sequences(oracle.make_params(*args, **kw), seed), gold, name + " restatement")
        assert_bitexact(harness.simulate(make_params(*args, **kw), seed)["seqs"], gold, name + " device code on the CPU")
