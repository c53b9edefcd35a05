"""Pins the CPU oracle (oracle/mcts_oracle.c) against fixtures generated from the
live reference (tests/golden/make_golden.py): RNG taps, post-shuffle sequences,
per-path (profit, drawdown) of all 20 strategies, and the summary rows -- all
bit-exact.  CPU only."""
import numpy as np
import pytest

from conftest import assert_bitexact
from oracle import oracle


def _params(meta):
    return oracle.make_params(*meta["args"], **meta["kwargs"])


def _case_names(kat):
    return sorted(kat["meta"].keys())


def test_mt19937_taps_match_numpy(rng_golden):
    for seed in (0, 1, 12345, 4294967295):
        d, s = oracle.rng_taps(seed, 700, 1000)
        assert_bitexact(d, rng_golden["rand/%d" % seed], "random_sample seed %d" % seed)
        assert_bitexact(s, rng_golden["shuffle/%d" % seed], "shuffle seed %d" % seed)


def test_seeded_runs_reproduce_reference_bit_for_bit(kat):
    """oracle(seed) == reference(seed): generators + shuffle + strategies + aggregates."""
    for name in _case_names(kat):
        meta = kat["meta"][name]
        out = oracle.run_all(_params(meta), meta["seed"], per_path=True, sequences=True)
        assert_bitexact(out["seqs"], kat[name + "/seqs"], name + " sequences")
        assert_bitexact(out["profit"], kat[name + "/profit"], name + " profit")
        assert_bitexact(out["dd"], kat[name + "/dd"], name + " drawdown")
        ids = kat[name + "/rank_ids"]
        assert_bitexact(out["summary"][ids - 1], kat[name + "/rows"], name + " summary rows")


def test_replay_entry_matches_reference_per_path(kat):
    """oracle.eval_paths on the reference's own sequences == the reference's per-path results."""
    for name in _case_names(kat):
        meta = kat["meta"][name]
        seqs = kat[name + "/seqs"]
        out = oracle.eval_paths(seqs, num_trades=meta["args"][3])
        assert_bitexact(out["profit"], kat[name + "/profit"], name + " profit")
        assert_bitexact(out["dd"], kat[name + "/dd"], name + " drawdown")
        ids = kat[name + "/rank_ids"]
        assert_bitexact(out["summary"][ids - 1], kat[name + "/rows"], name + " summary rows")


def test_full_size_c1_c2_reproduce_reference(full_golden):
    """BASELINE configs C1 (No Markov) and C2 (Markov-1) at their stated 1,000 paths x 1,000 trades: seeded oracle
    run == live reference (sequences, all 20 x 1,000 per-path results, ranked rows), and the replay entry agrees."""
    for name in sorted(full_golden["meta"]):
        meta = full_golden["meta"][name]
        out = oracle.run_all(_params(meta), meta["seed"], per_path=True, sequences=True)
        assert out["seqs"].shape == (1000, 1000)
        assert_bitexact(out["seqs"], full_golden[name + "/seqs"], name + " sequences")
        assert_bitexact(out["profit"], full_golden[name + "/profit"], name + " profit")
        assert_bitexact(out["dd"], full_golden[name + "/dd"], name + " drawdown")
        ids = full_golden[name + "/rank_ids"]
        assert_bitexact(out["summary"][ids - 1], full_golden[name + "/rows"], name + " summary rows")
        rep = oracle.eval_paths(full_golden[name + "/seqs"], num_trades=1000)
        assert_bitexact(rep["profit"], full_golden[name + "/profit"], name + " replay profit")
        assert_bitexact(rep["summary"][ids - 1], full_golden[name + "/rows"], name + " replay rows")


def test_survey_known_answers(kat):
    """SURVEY.md A.5 hex-exact anchors (independent of the fixture file)."""
    row = lambda name, sid: kat[name + "/rows"][list(kat[name + "/rank_ids"]).index(sid)]
    assert row("a5_dynamic", 1)[0] == float.fromhex("0x1.9433e66666665p+16")
    assert row("a5_dynamic", 1)[1] == float.fromhex("-0x1.1c79999999968p+10")
    assert row("a5_dynamic", 2)[0] == float.fromhex("0x1.6b0c2cccccccdp+17")
    assert row("a5_dynamic", 19)[1] == float.fromhex("-0x1.2721999999997p+11")
    assert row("a5_markov1", 1)[0] == 63825.0 and row("a5_markov2", 19)[1] == -3997.5
    assert row("a5_regime", 2)[0] == -29466.666666666668


def test_np_sum_restatement_matches_numpy():
    rng = np.random.default_rng(7)
    for n in list(range(1, 40)) + [63, 64, 65, 120, 127, 128, 129, 136, 255, 256, 257, 1000, 1001, 4095, 4104, 20000]:
        a = rng.normal(scale=1e3, size=n) + rng.integers(-3, 3, size=n) * 204.60000000000002
        assert oracle.np_sum(a) == float(np.sum(a)), n
        assert oracle.np_sum(a) / n == float(np.mean(a)), n


def test_tie_groups_of_the_reference(kat):
    """strategies inside a class produce identical numbers."""
    groups = [(1, 3), (2, 13), (4, 5, 14, 17), (6, 7), (9, 10, 15), (11, 12)]
    for name in _case_names(kat):
        p, d = kat[name + "/profit"], kat[name + "/dd"]
        for g in groups:
            for other in g[1:]:
                assert_bitexact(p[g[0] - 1], p[other - 1], "%s profit %s" % (name, g))
                assert_bitexact(d[g[0] - 1], d[other - 1], "%s dd %s" % (name, g))


def test_error_behaviour():
    with pytest.raises(ValueError):  # np.min of an empty list raises in the reference
        oracle.run_all(oracle.make_params(0.8, 1.0, 1.0, 10, 0, 5), 1)
    with pytest.raises(ValueError):  # empty sequence
        oracle.run_all(oracle.make_params(0.8, 1.0, 1.0, 0, 2, 2), 1)
    with pytest.raises(ValueError):  # choice(): negative probability
        oracle.run_all(oracle.make_params(1.3, 1.0, 1.0, 10, 2, 2), 1)
