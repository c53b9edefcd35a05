#!/usr/bin/env python
"""Build an experimental variant of libmcts.so with extra nvcc flags:
    python tools/build_variant.py TAG [-DMACRO=..] ...   ->  tools/variants/libmcts_TAG.so
Run it with MCTS_LIB=tools/variants/libmcts_TAG.so (the product always loads the in-package library)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_trade_sizing_b200 import build as B  # noqa: E402

tag, extra = sys.argv[1], sys.argv[2:]
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "variants")
B.LIB = os.path.join(root, "libmcts_%s.so" % tag)
B.OBJDIR = os.path.join(root, "obj_" + tag)
print(B.build_lib(force=True, extra=extra))
