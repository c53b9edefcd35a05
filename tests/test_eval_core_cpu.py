"""CPU unit tests of the DEVICE evaluator code (csrc/eval_core.cuh compiled with g++ by
tests/cpu_harness): the reduced two-outcome rules, the history masks, the NumPy-ordered profit
accumulation and the drawdown recurrence must reproduce the reference bit for bit.  These run
without a GPU; the same comparisons run through libmcts on the B200 in test_gpu_parity.py."""
import numpy as np

import harness
from conftest import assert_bitexact
from oracle import oracle


def test_rules_reproduce_reference_golden(kat):
    for name in sorted(kat["meta"]):
        for dedup in (False, True):
            p, d = harness.eval_values(kat[name + "/seqs"], dedup)
            assert_bitexact(p, kat[name + "/profit"], "%s profit dedup=%s" % (name, dedup))
            assert_bitexact(d, kat[name + "/dd"], "%s drawdown dedup=%s" % (name, dedup))


def test_random_sequences_vs_literal_closures():
    """oracle.eval_paths restates the 20 closures literally (streak counters, pause mode); the device
    code uses the reduced rules of SURVEY A.2.  Includes zeros, -0.0, non-integers, all trade counts
    around the pairwise-tree boundaries."""
    rng = np.random.default_rng(11)
    pool = np.array([204.60000000000002, 167.4, -240.9, 186.0, -219.0, 0.0, -0.0, 1e-3, -7.25, 3e7, -1e-9])
    for T in list(range(1, 20)) + [63, 64, 65, 120, 127, 128, 129, 135, 136, 137, 255, 256, 257, 264, 511, 512, 513, 1000, 1001, 1007, 1100, 2049, 4100]:
        P = 12 if T < 600 else 4
        vals = rng.choice(pool, size=(P, T))
        vals[0] = rng.normal(size=T) * 100
        vals[1] = np.where(rng.random(T) < 0.97, 186.0, -219.0)   # long win streaks
        vals[2] = np.where(rng.random(T) < 0.03, 186.0, -219.0)   # long loss streaks
        ref = oracle.eval_paths(vals)
        p, d = harness.eval_values(vals, False)
        assert_bitexact(p, ref["profit"], "T=%d profit" % T)
        assert_bitexact(d, ref["dd"], "T=%d drawdown" % T)
        p2, d2 = harness.eval_values(vals, True)
        assert_bitexact(p2, p, "T=%d dedup profit" % T)
        assert_bitexact(d2, d, "T=%d dedup drawdown" % T)


def test_all_256_outcome_patterns_of_length_8():
    """Exhaustive over win/loss patterns of 8 trades (plus lengths 1..7 as prefixes): every DFA state
    of every strategy is reached."""
    pats = np.array([[186.0 if (m >> i) & 1 else -219.0 for i in range(8)] for m in range(256)])
    for T in range(1, 9):
        vals = np.ascontiguousarray(pats[:, :T])
        ref = oracle.eval_paths(vals)
        p, d = harness.eval_values(vals, False)
        assert_bitexact(p, ref["profit"], "T=%d" % T)
        assert_bitexact(d, ref["dd"], "T=%d" % T)
    pats12 = np.array([[204.60000000000002 if (m >> i) & 1 else -240.9 for i in range(12)] for m in range(4096)])
    ref = oracle.eval_paths(pats12)
    p, d = harness.eval_values(pats12, False)
    assert_bitexact(p, ref["profit"], "12-trade patterns")
    assert_bitexact(d, ref["dd"], "12-trade patterns")
