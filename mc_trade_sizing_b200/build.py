"""Builds libmcts.so (the CUDA core + C ABI) in-tree with nvcc for sm_100a.

    python -m mc_trade_sizing_b200.build [--force]

The shared object lands next to this file so that it travels to the GPU box with the repo
snapshot; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmcts.so")
SOURCES = ["mcts.cu", "paths_full.cu", "paths_dedup.cu", "paths_int.cu", "paths_curves.cu"]
HEADERS = ["stager.hpp", "kernels.cuh", "aux_kernels.cuh", "paths_launch.cuh", "eval_core.cuh", "gen_core.cuh", "pairwise.cuh", "planner.hpp", "spec.hpp",
           os.path.join("..", "..", "include", "mcts.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-fmad=false",            # equity + value*size must round the product first (no FMA contraction)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
]
LINK_FLAGS = ["-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
OBJDIR = os.path.join(HERE, "build")


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libmcts.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(args):
    src, obj, extra, verbose = args
    cmd = [nvcc()] + NVCC_FLAGS + list(extra) + (["-Xptxas=-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build_lib(force: bool = False, verbose: bool = False, extra=()) -> str:
    """Compile every translation unit for sm_100a (in parallel) and link libmcts.so."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJDIR, exist_ok=True)
    jobs = [(os.path.join(CSRC, s), os.path.join(OBJDIR, s.replace(".cu", ".o")), extra, verbose) for s in SOURCES]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
        results = list(ex.map(_compile, jobs))
    for src, rc, log in results:
        if rc != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, log))
        if verbose:
            sys.stderr.write(log)
    cmd = [nvcc()] + LINK_FLAGS + ["-o", LIB] + [j[1] for j in jobs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
