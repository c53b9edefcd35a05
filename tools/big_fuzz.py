import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import fuzz_cases
from mc_trade_sizing_b200 import Engine, make_params, _ffi
from oracle import oracle
eng = Engine(0)
ok = err = unsup = 0
for seed in (1, 2, 3, 4, 5, 6):
    for args, kw in fuzz_cases.random_cases(500, seed):
        prm = make_params(*args, **kw)
        try:
            ref = oracle.devgen_sequences(oracle.make_params(*args, **kw), 99 + seed)
        except ValueError:
            ref = None
        if ref is None or ref.shape[1] == 0:
            try:
                eng.simulate(prm, seed=99 + seed); raise SystemExit("expected error %s %s" % (args, kw))
            except ValueError:
                err += 1
            continue
        opts = eng.options(dedup=bool(ok & 1), int_fast_path=bool(ok & 2), chunk_paths=128 if ok & 4 else 0)
        try:
            out = eng.simulate(prm, seed=99 + seed, per_path=True, options=opts)
        except _ffi.MctsError as e:
            assert e.code == _ffi.EUNSUPPORTED; unsup += 1; continue
        exp = oracle.eval_paths(ref, num_trades=args[3])
        for k in ("profit", "dd"):
            if not (out[k].view(np.uint64) == exp[k].view(np.uint64)).all():
                raise SystemExit("MISMATCH %s %s %s" % (k, args, kw))
        if not (eng.sequences(prm, seed=99 + seed).view(np.uint64) == ref.view(np.uint64)).all():
            raise SystemExit("SEQ MISMATCH %s %s" % (args, kw))
        ok += 1
print("big fuzz: ok", ok, "errors-as-expected", err, "unsupported", unsup)
