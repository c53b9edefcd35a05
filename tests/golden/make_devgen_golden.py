#!/usr/bin/env python
"""Known answers of the device generator's published stream (MCTS_GENERATOR_VERSION 3):

    python tests/golden/make_devgen_golden.py        ->  tests/golden/devgen_v3.npz

The sequences come from the INDEPENDENT plain-C restatement of the spec (oracle/devgen_ref.c, test infrastructure), not from
the product; tests/test_generator_cpu.py (g++ build of the device headers) and tests/test_gpu_round2.py (the kernels) must
reproduce them bit for bit.  The file pins the stream: a change of the generator that alters any of these numbers needs a
new MCTS_GENERATOR_VERSION and a new fixture (results are documented as a pure function of arguments, seed and that version).
Needs neither /root/reference nor a GPU.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle  # noqa: E402
from test_generator_cpu import URN_EDGE_CASES  # noqa: E402

SEED = 20261020
CASES = list(URN_EDGE_CASES) + [
    ("markov1", (0.82, 186.0, 219.0, 96, 4, 3), dict(use_markov=True, p_win_after_win=0.75, p_win_after_loss=0.60)),
    ("markov2", (0.82, 186.0, 219.0, 97, 4, 3), dict(use_markov2=True, p_win_ww=0.85, p_win_wl=0.65, p_win_lw=0.55, p_win_ll=0.30)),
    ("c5_dynamic_1000", (0.82, 186.0, 219.0, 1000, 2, 2), {}),
]

if __name__ == "__main__":
    out = {}
    for name, args, kw in CASES:
        seqs = oracle.devgen_sequences(oracle.make_params(*args, **kw), SEED)
        vals, codes = np.unique(seqs, return_inverse=True)
        assert len(vals) < 256
        out[name + "__values"] = vals
        out[name + "__codes"] = codes.reshape(seqs.shape).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "devgen_v3.npz"), **out)
    print("wrote devgen_v3.npz:", {k: v.shape for k, v in out.items() if k.endswith("__codes")})
