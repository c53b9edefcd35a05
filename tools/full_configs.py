#!/usr/bin/env python
"""The five configurations of BASELINE.json at FULL size on one B200 (or sharded over the ranks of a
torchrun launch): wall time through the public API, trade-steps/s and the strategy-1 mean against the
analytic expectation.

    python tools/full_configs.py [--skip-c5]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/full_configs.py
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_trade_sizing_b200 import Engine, make_params  # noqa: E402
from mc_trade_sizing_b200 import distributed as mdist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--skip-c5", action="store_true")
ap.add_argument("--c5-sims", type=int, default=50_000_000)
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
device = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)

REG = [dict(length=300, hit_rate=0.9, avg_win=200, avg_loss=100), dict(length=200, hit_rate=0.5, avg_win=100, avg_loss=100),
       dict(length=500, hit_rate=0.2, avg_win=100, avg_loss=200)]
CONFIGS = [
    ("C1 default yaml (No Markov), 50 x 20 = 1,000 paths", (0.82, 186.0, 219.0, 1000, 50, 20), {}, 103063.2),
    ("C2 Markov-1 0.75/0.60, 1,000 paths", (0.82, 186.0, 219.0, 1000, 50, 20), dict(use_markov=True, p_win_after_win=0.75, p_win_after_loss=0.60), 66936.7),
    ("C3 Markov-2 0.85/0.65/0.55/0.30, 1e6 paths", (0.82, 186.0, 219.0, 1000, 50000, 20),
     dict(use_markov2=True, p_win_ww=0.85, p_win_wl=0.65, p_win_lw=0.55, p_win_ll=0.30), 57802.5),
    ("C4 regime switching (README), 1e7 paths", (0.82, 186.0, 219.0, 1000, 500000, 20), dict(use_regime=True, regimes=REG), -19000.0),
]
if not a.skip_c5:
    CONFIGS.append(("C5 scale sweep (No Markov), %.3g paths" % (a.c5_sims * 20), (0.82, 186.0, 219.0, 1000, a.c5_sims, 20), {}, 103063.2))

eng = Engine(local)
eng.simulate(make_params(0.82, 186.0, 219.0, 1000, 1000, 20))  # warm-up: context, buffers
if world > 1:  # ... and the communicator: the first NCCL collective of a process group takes ~1 s (it used to land in C1's wall time)
    wp = make_params(0.82, 186.0, 219.0, 1000, 4096 * world, 20)
    wch = mdist.resolve_chunk(4096 * world * 20, 0, world)
    mdist.simulate_sharded(wp, 1, lambda lo, hi: eng.simulate_chunks(wp, 1, lo, hi, options=eng.options(chunk_paths=wch)), rank, world, chunk_paths=wch, device=device)
if rank == 0:
    print("%d GPU(s)" % world)
for name, args, kw, expect in CONFIGS:
    eng.simulate(make_params(*(args[:4] + (8, args[5])), **kw))  # 160 paths of the same mode: loads this mode's kernels (lazy module loading)
    prm = make_params(*args, **kw)
    P = args[4] * args[5]
    t0 = time.perf_counter()
    if world == 1:
        rows = eng.simulate(prm, seed=2026)["summary"]
    else:
        ch = mdist.resolve_chunk(P, 0, world)   # every rank gets chunks, also for the 1,000-path configs' single chunk
        _, rows = mdist.simulate_sharded(prm, 2026, lambda lo, hi: eng.simulate_chunks(prm, 2026, lo, hi, options=eng.options(chunk_paths=ch)), rank, world,
                                         chunk_paths=ch, device=device)
    wall = time.perf_counter() - t0
    if rank == 0:
        print("%-52s wall %9.3f ms  %.3e trade-steps/s  avg_profit[1] = %.3f (analytic %.1f)  avg_dd[1] = %.3f"
              % (name, wall * 1e3, P * 1000 * 20 / wall, rows[0][0], expect, rows[0][1]), flush=True)

if world == 1:
    # C4's equity-curve export at full size: 1e7 paths x 20 x 1000 x 8 B = 1.6 TB, streamed through one 21 GB device buffer
    # (131,072 paths per batch, overwritten by the next batch; draining to the host would be PCIe-bound)
    prm = make_params(0.82, 186.0, 219.0, 1000, 500000, 20, use_regime=True, regimes=REG)
    batch, total = 1 << 17, 10_000_000
    dev = eng.device_alloc(20 * 500 * batch * 16)
    eng.export_curves(prm, 1, 0, batch, to_host=False, dev=dev)
    t0 = time.perf_counter()
    ms = 0.0
    for first in range(0, total, batch):
        n = min(batch, total - first)
        eng.export_curves(prm, 2026, first, n, to_host=False, dev=dev)
        ms += eng.timing()["paths_ms"]
    wall = time.perf_counter() - t0
    eng.device_free(dev)
    print("C4 curve export, 1e7 paths (1.6 TB through a 21 GB device buffer): wall %.1f ms, kernels %.1f ms = %.0f GB/s, %.3e trade-steps/s"
          % (wall * 1e3, ms, total * 20000 * 8 / ms / 1e6, total * 20000 / wall))
if world > 1:
    dist.destroy_process_group()
