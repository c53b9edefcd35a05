#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

The reference module (src/trading_models.py) is imported read-only, its global
NumPy RandomState is seeded by this harness (the reference never seeds), and
`strategy_static` / `strategy_dynamic` are wrapped as module globals to capture
the post-shuffle sequence and the per-path (profit, drawdown) of every strategy
without touching reference source.

    python tests/golden/make_golden.py --full     # full.npz only (~1 min)

Outputs (committed):
    full.npz    BASELINE configs C1 and C2 at their stated size (50 simulations x 20 shuffles = 1,000 paths x
                1,000 trades) from the live reference: sequences (as a value table + uint8 codes), per-path
                results of all 20 strategies, ranked summary rows
    kat.npz     seeded known-answer runs: sequences, per-path results, summaries
    stats.npz   larger per-path samples per mode for statistical tests of the
                device RNG path
    rng.npz     raw NumPy legacy RandomState taps (random_sample, shuffle)
"""
import ast
import importlib.util
import inspect
import json
import os
import sys

import numpy as np

REF = os.environ.get("MCTS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_trading_models", os.path.join(REF, "src", "trading_models.py"))
    tm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tm)
    return tm


def descriptions_by_id(tm):
    """The id -> description dict literal inside run_all_strategies (:451-472)."""
    tree = ast.parse(inspect.getsource(tm.run_all_strategies))
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "descriptions":
            return ast.literal_eval(node.value)
    raise RuntimeError("descriptions dict not found")


def capture_run(tm, seed, args, kwargs):
    seqs, per = [], []
    orig_static, orig_dynamic = tm.strategy_static, tm.strategy_dynamic

    def static(results):
        seqs.append(np.array(results, dtype=np.float64))
        out = orig_static(results)
        per.append((float(out[0]), float(out[1])))
        return out

    def dynamic(results, cond):
        out = orig_dynamic(results, cond)
        per.append((float(out[0]), float(out[1])))
        return out

    tm.strategy_static, tm.strategy_dynamic = static, dynamic
    try:
        np.random.seed(seed)
        rows = tm.run_all_strategies(*args, **kwargs)
    finally:
        tm.strategy_static, tm.strategy_dynamic = orig_static, orig_dynamic
    P = len(seqs)
    per = np.array(per, dtype=np.float64).reshape(P, 20, 2)
    return {
        "seqs": np.stack(seqs) if P else np.zeros((0, 0)),
        "profit": np.ascontiguousarray(per[:, :, 0].T),
        "dd": np.ascontiguousarray(per[:, :, 1].T),
        "rows": rows,
    }


README_REGIMES = [
    {"length": 300, "hit_rate": 0.9, "avg_win": 200, "avg_loss": 100},
    {"length": 200, "hit_rate": 0.5, "avg_win": 100, "avg_loss": 100},
    {"length": 500, "hit_rate": 0.2, "avg_win": 100, "avg_loss": 200},
]
FLOAT_REGIMES = [
    {"length": 120, "hit_rate": 0.7, "avg_win": 150.5, "avg_loss": 99.3},
    {"length": 80, "hit_rate": 0.35, "avg_win": 61.7, "avg_loss": 140.1},
]
MODE_KW = {
    "dynamic": {},
    "markov1": dict(use_markov=True, p_win_after_win=0.75, p_win_after_loss=0.60),
    "markov2": dict(use_markov2=True, p_win_ww=0.85, p_win_wl=0.65, p_win_lw=0.55, p_win_ll=0.30),
    "regime": dict(use_regime=True, regimes=README_REGIMES),
}


def kat_cases():
    cases = []
    base = (0.82, 186.0, 219.0)
    for mode, kw in MODE_KW.items():  # SURVEY A.5 known answers: 3 sims x 2 shuffles, seed 12345
        cases.append(("a5_%s" % mode, 12345, base + (1000, 3, 2), kw))
    for T in (1, 2, 3, 7, 8, 9, 15, 16, 17, 127, 128, 129, 255, 256, 257, 300):
        for mode, kw in MODE_KW.items():
            kw2 = dict(kw)
            if mode == "regime":
                kw2["regimes"] = None  # defaults scale with T (:108-112)
            cases.append(("T%d_%s" % (T, mode), 1000 + T, base + (T, 2, 3), kw2))
    for h in (0.72, 0.7699999999999999, 0.82):  # the three hit rates dps.py:187-191 passes
        cases.append(("dps_h%.2f" % h, 77, (h, 186.0, 219.0, 1000, 4, 5), {}))
    cases.append(("nonint_markov1", 5, (0.6, 186.3, 219.7, 500, 3, 3), MODE_KW["markov1"]))
    cases.append(("nonint_markov2", 6, (0.6, 186.3, 219.7, 500, 3, 3), MODE_KW["markov2"]))
    cases.append(("regime_float", 7, base + (200, 3, 3), dict(use_regime=True, regimes=FLOAT_REGIMES)))
    cases.append(("regime_short", 8, base + (1200, 2, 3), dict(use_regime=True, regimes=README_REGIMES)))
    cases.append(("regime_trunc", 9, base + (450, 2, 3), dict(use_regime=True, regimes=README_REGIMES)))
    cases.append(("regime_default", 10, base + (1000, 2, 3), dict(use_regime=True)))
    cases.append(("lowhit_dynamic", 11, (0.2, 50.0, 75.5, 333, 3, 4), {}))
    cases.append(("precedence", 12, base + (64, 2, 2), dict(use_markov=True, use_markov2=True)))
    return cases


def main():
    tm = load_reference()
    desc = descriptions_by_id(tm)
    desc_to_id = {v: k for k, v in desc.items()}
    out = {"descriptions": np.array([desc[i] for i in range(1, 21)])}
    meta = {}
    for name, seed, args, kw in kat_cases():
        r = capture_run(tm, seed, args, kw)
        out[name + "/seqs"] = r["seqs"]
        out[name + "/profit"] = r["profit"]
        out[name + "/dd"] = r["dd"]
        out[name + "/rank_ids"] = np.array([desc_to_id[row[0]] for row in r["rows"]], dtype=np.int32)
        out[name + "/rows"] = np.array([[float(x) for x in row[1:]] for row in r["rows"]], dtype=np.float64)
        meta[name] = {"seed": seed, "args": list(args), "kwargs": kw}
        print("kat", name, r["seqs"].shape)
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **out)

    stats, smeta = {}, {}
    for mode, kw in MODE_KW.items():
        args = (0.82, 186.0, 219.0, 1000, 300, 4)
        r = capture_run(tm, 2024, args, kw)
        stats[mode + "/profit"] = r["profit"]
        stats[mode + "/dd"] = r["dd"]
        smeta[mode] = {"seed": 2024, "args": list(args), "kwargs": kw}
        print("stats", mode, r["profit"].shape)
    stats["meta"] = np.array(json.dumps(smeta))
    np.savez_compressed(os.path.join(HERE, "stats.npz"), **stats)

    rng = {}
    for seed in (0, 1, 12345, 4294967295):
        np.random.seed(seed)
        rng["rand/%d" % seed] = np.array([np.random.rand() for _ in range(700)])
        x = np.arange(1000, dtype=np.float64)
        np.random.shuffle(x)
        rng["shuffle/%d" % seed] = x
    np.savez_compressed(os.path.join(HERE, "rng.npz"), **rng)
    print("numpy", np.__version__, "python", sys.version.split()[0])


def make_full_golden():
    """C1 and C2 of BASELINE.json at full size, from the live reference."""
    tm = load_reference()
    desc = descriptions_by_id(tm)
    desc_to_id = {v: k for k, v in desc.items()}
    out, meta = {}, {}
    for name, seed, args, kw in (("c1_full", 20261, (0.82, 186.0, 219.0, 1000, 50, 20), MODE_KW["dynamic"]),
                                 ("c2_full", 20262, (0.82, 186.0, 219.0, 1000, 50, 20), MODE_KW["markov1"])):
        r = capture_run(tm, seed, args, kw)
        table, codes = np.unique(r["seqs"], return_inverse=True)   # <= 6 distinct trade values per run
        assert table.size <= 255 and (table[codes.reshape(r["seqs"].shape)].view(np.uint64) == r["seqs"].view(np.uint64)).all()
        out[name + "/values"] = table
        out[name + "/codes"] = codes.reshape(r["seqs"].shape).astype(np.uint8)
        out[name + "/profit"] = r["profit"]
        out[name + "/dd"] = r["dd"]
        out[name + "/rank_ids"] = np.array([desc_to_id[row[0]] for row in r["rows"]], dtype=np.int32)
        out[name + "/rows"] = np.array([[float(x) for x in row[1:]] for row in r["rows"]], dtype=np.float64)
        meta[name] = {"seed": seed, "args": list(args), "kwargs": kw}
        print("full", name, r["seqs"].shape, table)
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "full.npz"), **out)


if __name__ == "__main__" and "--full" in sys.argv:
    make_full_golden()
elif __name__ == "__main__" and "--cli" not in sys.argv:
    main()
    make_full_golden()


def make_cli_golden():
    """stdout of the reference's own main() (src/trading_models.py:532-671) with run_all_strategies
    replaced by a stub that returns the rows captured in kat.npz -- pins the CLI text format
    (settings block, break-even line, table, rule after row 3) independently of the RNG."""
    import contextlib
    import io
    tm = load_reference()
    desc = descriptions_by_id(tm)
    kat = np.load(os.path.join(HERE, "kat.npz"))
    cases = {
        "a5_dynamic": ["--hit_rate", "0.82", "--avg_win", "186.0", "--avg_loss", "219.0", "--num_simulations", "3",
                       "--num_trades", "1000", "--num_mc_shuffles", "2"],
        "a5_markov1": ["--hit_rate", "0.7699999999999999", "--avg_win", "186.0", "--avg_loss", "219.0", "--num_simulations", "3",
                       "--num_trades", "1000", "--num_mc_shuffles", "2", "--use_markov", "--p_win_after_win", "0.75",
                       "--p_win_after_loss", "0.6"],
        "a5_markov2": ["--hit_rate", "0.72", "--avg_win", "186.0", "--avg_loss", "219.0", "--num_simulations", "3",
                       "--num_trades", "1000", "--num_mc_shuffles", "2", "--use_markov2", "--p_win_ww", "0.85", "--p_win_wl", "0.65",
                       "--p_win_lw", "0.55", "--p_win_ll", "0.3"],
        "a5_regime": ["--hit_rate", "0.82", "--avg_win", "186.0", "--avg_loss", "219.0", "--num_simulations", "3",
                      "--num_trades", "1000", "--num_mc_shuffles", "2", "--use_regime", "--regimes", json.dumps(README_REGIMES)],
    }
    out = {}
    sys.modules["colorama"] = None  # the reference prints the plain variant when colorama is absent (:647-671)
    for name, argv in cases.items():
        ids, rows = kat[name + "/rank_ids"], kat[name + "/rows"]
        table = [(desc[int(i)],) + tuple(np.float64(x) for x in r) for i, r in zip(ids, rows)]
        orig = tm.run_all_strategies
        tm.run_all_strategies = lambda *a, **k: table
        buf = io.StringIO()
        old_argv = sys.argv
        sys.argv = ["trading_models.py"] + argv
        try:
            with contextlib.redirect_stdout(buf):
                tm.main()
        finally:
            sys.argv = old_argv
            tm.run_all_strategies = orig
        out[name] = {"argv": argv, "stdout": buf.getvalue()}
        print("cli", name, len(buf.getvalue()), "bytes")
    with open(os.path.join(HERE, "cli.json"), "w") as f:
        json.dump(out, f, indent=1, ensure_ascii=False)


if __name__ == "__main__" and "--cli" in sys.argv:
    make_cli_golden()
