#!/usr/bin/env python
"""bench.py -- throughput of the Monte Carlo path-simulation hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): trade-steps/s = paths x trades x 20 strategies / second.
Workload: config C5 (scale sweep) -- the "No Markov" three-phase generator of dps_config.yaml
(hit_rate 0.82, avg_win 186, avg_loss 219), 1000 trades, 20 shuffles per simulation, all 20
strategy state machines evaluated -- cut into per-step batches of SIMS_PER_GPU_STEP simulations per
GPU (weak scaling: the full 1e9-path sweep is ~240 such batches per GPU on one GPU).  One step = one
complete pass of the hot path over one batch: base generation, per-path shuffle, 20 strategies,
exact-order aggregation (+ the record all_gather and tree combine when N > 1).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, NUM_SHUFFLES = 0.82, 186.0, 219.0, 1000, 20
SIMS_PER_GPU_STEP = 419432          # x 20 shuffles = 8,388,640 paths = 1.68e11 trade-steps per GPU per step
CHUNK_PATHS = 1 << 22          # the library default: two 4.19 M-path chunks per GPU per step, one per stream lane
METRIC = "trade-steps/sec (paths x trades x strategies)"
UNIT = "trade-steps/s"
FP64_OPS_PER_TRADE_STEP = 6          # mul, add(equity), max, sub, min, add(profit)
TRAFFIC_PER_LAUNCH = 645.3e6         # dram bytes read + written by one k_sim_paths launch: a CONSTANT from the committed ncu capture
TRAFFIC_SOURCE = "ncu constant: dram__bytes_read.sum + dram__bytes_write.sum of one launch in profiles/r02_ncu_bench_sim_paths.txt (not re-measured by bench.py)"
C5_FULL_SIMS = 50_000_000            # x 20 shuffles = 1e9 paths: BASELINE.json config 5 at its stated size
STRONG_SIMS = 3_355_456              # x 20 = 67,109,120 paths: the FIXED problem of the strong-scaling line (8 x the per-GPU batch)
SHA_SIMS, SHA_SHUFFLES = 1 << 20, 16  # 2^24 paths: the fixed run whose summary hash must not depend on the number of GPUs


def workload_config(n_gpus):
    return {
        "workload": "C5 scale-sweep batch: 'No Markov' 3-phase generator hit_rate=0.82 avg_win=186 avg_loss=219, "
                    "1000 trades, 20 shuffles/sim, 20 strategies (all 20 state machines evaluated), "
                    "%d sims (%d paths) per GPU per step" % (SIMS_PER_GPU_STEP, SIMS_PER_GPU_STEP * NUM_SHUFFLES),
        "paths_per_step": SIMS_PER_GPU_STEP * NUM_SHUFFLES * n_gpus,
        "trades": NUM_TRADES, "strategies": 20, "generator": "dynamic", "rng": "philox4x32-10, new seed every step",
        "parallelism": "paths sharded by pairwise-tree node, %d GPU(s), one record all_gather per step" % n_gpus,
        "l2": "no reusable inputs: every step regenerates its outcomes on device from a new seed; the per-step "
              "per-path result buffer (40 x 8 B x %d paths per chunk = %d MB) exceeds the 126 MB L2" % (CHUNK_PATHS, 320 * CHUNK_PATHS >> 20),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(target_seconds=12.0, cores=None):
    """The reference's algorithm on the host cores: oracle/mcts_oracle.c (a C restatement pinned
    bit-exact to the live Python reference; the Python original is ~100x slower per core and cannot
    travel to the GPU box).  One independent run_all_strategies per thread."""
    from oracle import oracle
    cores = cores or host_cores()
    prm = oracle.make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, 5, NUM_SHUFFLES)
    t = oracle.bench(prm, 1, cores)                         # calibration: 100 paths per thread
    per_sim = max(t / 5, 1e-6)
    sims = max(5, int(target_seconds / per_sim))
    prm = oracle.make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, sims, NUM_SHUFFLES)
    secs = oracle.bench(prm, 1000, cores)
    steps = cores * sims * NUM_SHUFFLES * NUM_TRADES * 20
    return {"value": steps / secs, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d threads x %d sims x %d shuffles x %d trades x 20 strategies of the same C5 parameters "
                      "(oracle/mcts_oracle.c incl. MT19937 + Fisher-Yates), %.1f s wall" % (cores, sims, NUM_SHUFFLES, NUM_TRADES, secs),
            "seconds": secs, "trade_steps": steps}


def python_reference_rates(cores=None):
    """The REAL Python reference (an unmodified copy under baseline/_ref, see baseline/install_ref.py) timed on this
    box's host cores the way BASELINE.md section 3 prescribes: in-process run_all_strategies, NumPy seeded per worker,
    (1) one core per generator mode at 10 x 20 = 200 paths, (2) one worker per host core, (3) three concurrent workers
    (what dps.py:268 achieves).  Each worker is a separate process (baseline/ref_worker.py); a group's aggregate rate is
    its total trade-steps over the slowest worker's time inside the call."""
    worker = os.path.join(ROOT, "baseline", "ref_worker.py")
    if not os.path.exists(os.path.join(ROOT, "baseline", "_ref", "src", "trading_models.py")):
        return {"unavailable": "baseline/_ref is not populated (python baseline/install_ref.py needs /root/reference, which exists only in the build container)"}
    cores = cores or host_cores()
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")

    def group(jobs):
        start_at = time.time() + 2.0 + 0.02 * len(jobs)        # every worker finishes its imports first
        procs = [subprocess.Popen([sys.executable, worker, mode, str(sims), str(sh), str(seed), repr(start_at)], stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL, text=True, env=env) for mode, sims, sh, seed in jobs]
        outs = []
        for pr in procs:
            out, _ = pr.communicate(timeout=600)
            if pr.returncode == 0 and out.strip():
                outs.append(json.loads(out.strip().splitlines()[-1]))
        return outs

    try:
        t0 = time.time()
        one = group([(m, 10, 20, 12345) for m in ("shuffle", "markov1", "markov2", "regime")])
        allc = group([("shuffle", 5, 20, 100 + i) for i in range(cores)])
        three = group([("shuffle", 5, 20, 200 + i) for i in range(3)])
        agg = lambda outs: sum(o["trade_steps"] for o in outs) / max(o["seconds"] for o in outs) if outs else None
        return {
            "kind": "reference", "unit": UNIT, "impl": "unmodified src/trading_models.py run_all_strategies, in process, np.random.seed per worker",
            "one_core": {o["mode"]: o["trade_steps"] / o["seconds"] for o in one},
            "one_core_sample": "10 sims x 20 shuffles x 1000 trades x 20 strategies per mode (parameters of configs C1-C4), one process per mode",
            "all_cores": {"value": agg(allc), "cores": len(allc), "sample": "%d workers x 5 sims x 20 shuffles x 1000 trades x 20 strategies (C5 parameters)" % len(allc)},
            "three_procs": {"value": agg(three), "cores": len(three), "sample": "3 workers x 5 x 20 paths: the concurrency dps.py:268 uses"},
            "avg_profit_strategy1": one[0]["avg_profit_strategy1"] if one else None,
            "seconds": time.time() - t0,
        }
    except Exception as e:  # a baseline must never cost the bench line
        return {"unavailable": "python reference run failed: %r" % (e,)}


def dps_overlay_walls():
    """The reference's UNCHANGED orchestrator (`python dps.py`: 12 CLI subprocesses, three at a time, its shipped
    dps_config.yaml = 12 x 4,000 paths x 1,000 trades x 20 = 9.6e8 trade-steps, scrape, sinks) against the drop-in through
    the overlay of INTEGRATION.md section 3 -- directly (12 CUDA contexts) and through the engine daemon (MCTS_SERVER=1).
    Uses the test helper tests/dps_overlay.py and the stand-ins for openpyxl / influxdb_client / flask under tests/stubs."""
    import shutil
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import dps_overlay
    except Exception as e:
        return {"unavailable": "tests/dps_overlay.py not importable: %r" % (e,)}
    ref = dps_overlay.reference_dir()
    if ref is None:
        return {"unavailable": "no reference copy (baseline/_ref) on this box"}
    out = {"runs": 12, "trade_steps": 960000000, "reference_wall_s_estimate": 9.6e8 / 5.0e6,
           "reference_estimate_note": "9.6e8 trade-steps at the three-process rate of cpu_baseline.python_reference (~5e6/s): about 3-5 min"}
    tmp = tempfile.mkdtemp(prefix="mcts_dps_")
    try:
        for name, served in (("direct", False), ("daemon", True)):
            ov = dps_overlay.build_overlay(os.path.join(tmp, name), ref)
            sdir = os.path.join(tmp, name + "-server")
            env = dict(MCTS_SERVER="1", MCTS_SERVER_DIR=sdir, MCTS_SERVER_SOCKET=os.path.join(sdir, "server.sock"), MCTS_SERVER_IDLE="30") if served else {"MCTS_SERVER": "0"}
            t0 = time.perf_counter()
            p = dps_overlay.run_dps(ov, extra_env=env, timeout=600)
            wall = time.perf_counter() - t0
            rows = len(dps_overlay.check_outputs(ov)) if p.returncode == 0 else 0
            out[name] = {"wall_s": wall, "returncode": p.returncode, "scraped_rows": rows}
            if served:
                subprocess.run([sys.executable, "-m", "mc_trade_sizing_b200.server", "--stop"], env=dict(os.environ, PYTHONPATH=ROOT, **env), cwd=ROOT,
                               capture_output=True, timeout=30)
    except Exception as e:
        out["error"] = repr(e)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU path on the box's host cores (all threads), same
    metric / config.  Rank 0 alone runs it."""
    if rank != 0:
        return
    cores = host_cores()
    per_step = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_reference_rate(target_seconds=4.0, cores=cores)
        if i >= args.warmup:
            per_step.append(base)
    secs = sum(b["seconds"] for b in per_step)
    steps = sum(b["trade_steps"] for b in per_step)
    value = steps / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": per_step[-1]["sample"],
                         "python_reference": python_reference_rates(cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python (no compilable sources): the VALUE of this arm is oracle/mcts_oracle.c, the C restatement "
                "pinned bit-exact to it, on all host threads -- about 100x faster per core than the Python original, i.e. the "
                "conservative denominator; the Python original itself is timed beside it under cpu_baseline.python_reference. "
                "Each step is a bounded sample of the C5 workload (config.sample_paths_per_step), not the full batch",
    }
    line["config"]["sample_paths_per_step"] = int(per_step[-1]["trade_steps"] // (NUM_TRADES * 20))
    print(json.dumps(line), flush=True)


def np_bytes(rows):
    import numpy as np
    return np.ascontiguousarray(np.asarray(rows, dtype=np.float64)).tobytes()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip dedup / curve-export side measurements")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # stdout carries the ONE JSON line: anything native libraries print meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from mc_trade_sizing_b200 import Engine, make_params
    from mc_trade_sizing_b200 import distributed as mdist
    from mc_trade_sizing_b200 import trading_models as tm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libmcts has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        host_group = dist.new_group(backend="gloo")   # CPU-side waits: ranks that idle must not spin on their GPU
    n_gpus = world

    eng = Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=device)
    peaks = eng.measure_peaks()

    n_sims = SIMS_PER_GPU_STEP * n_gpus
    prm = make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, n_sims, NUM_SHUFFLES)
    P = n_sims * NUM_SHUFFLES
    opt = eng.options(chunk_paths=CHUNK_PATHS)
    steps_per_step_global = P * NUM_TRADES * 20

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stat = {"launches": 0, "paths_ms": 0.0, "paths_launches": 0, "local_trade_steps": 0, "total_ms": 0.0}

    def one_step(seed, record):
        def compute(lo, hi):
            recs, n = eng.simulate_chunks(prm, seed, lo, hi, options=opt)
            if record and hi > lo:
                t = eng.timing()
                stat["launches"] += t["launches"]
                stat["paths_ms"] += t["paths_ms"]
                stat["paths_launches"] += t["paths_launches"]
                stat["local_trade_steps"] += t["trade_steps"]
                stat["total_ms"] += t["total_ms"]
            return recs, n
        return mdist.simulate_sharded(prm, seed, compute, rank, world, chunk_paths=CHUNK_PATHS, device=device if world > 1 else None)

    for w in range(args.warmup):
        one_step(1000 + w, False)

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t0 = time.perf_counter()
    last = None
    for s in range(args.steps):
        last = one_step(5000 + s, True)
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    tmax = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(tmax[0]), float(tmax[1])
    value = steps_per_step_global * args.steps / (dev_ms * 1e-3)

    # e2e: the public API a user calls (run_all_strategies / its sharded form), host buffers in and out
    def e2e_step(seed):
        if world == 1:
            return tm.run_all_strategies(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, n_sims, NUM_SHUFFLES, seed=seed,
                                         device=local_rank, chunk_paths=CHUNK_PATHS)
        return mdist.run_all_strategies_sharded(eng, prm, seed, rank, world, chunk_paths=CHUNK_PATHS, device=device)
    for w in range(args.warmup):   # the module-level engine of run_all_strategies allocates its device buffers on first use
        e2e_step(8000 + w)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        rows = e2e_step(9000 + s)
    barrier()
    e2e_wall = time.perf_counter() - t0
    t_e2e = (tm.get_engine(local_rank) if world == 1 else eng).timing()
    e2e_t = torch.tensor([e2e_wall], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = steps_per_step_global * args.steps / float(e2e_t[0])

    # dominant-kernel timing for the roofline: the same workload with chunks serialised on one stream, so
    # that the CUDA events around each k_sim_paths launch do not overlap a neighbouring launch
    kstat = {"paths_ms": 0.0, "paths_launches": 0, "trade_steps": 0, "total_ms": 0.0}
    opt_serial = eng.options(chunk_paths=CHUNK_PATHS, serial_chunks=True)
    lo, hi = Engine.plan_shards(P, CHUNK_PATHS, world)[rank]
    for s in range(3):
        eng.simulate_chunks(prm, 7000 + s, lo, hi, options=opt_serial)
        if s == 0:
            continue
        t = eng.timing()
        for k in kstat:
            kstat[k] += t[k]

    extras = {}
    if not args.no_extras and rank == 0:
        # (a) the 11 distinct behaviours evaluated once and fanned out (identical results)
        small = make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, SIMS_PER_GPU_STEP // 2, NUM_SHUFFLES)
        o2 = eng.options(chunk_paths=CHUNK_PATHS, dedup=True)
        eng.simulate(small, seed=1, options=o2)
        eng.simulate(small, seed=2, options=o2)
        t = eng.timing()
        extras["dedup11_value"] = t["trade_steps"] / (t["total_ms"] * 1e-3)
        # (a2) the other generator modes of BASELINE.json's configs (integer-valued -> exact int32 evaluator)
        kws = {"markov1_c2": dict(use_markov=True, p_win_after_win=0.75, p_win_after_loss=0.60),
               "markov2_c3": dict(use_markov2=True, p_win_ww=0.85, p_win_wl=0.65, p_win_lw=0.55, p_win_ll=0.30),
               "regime_c4": dict(use_regime=True, regimes=[dict(length=300, hit_rate=0.9, avg_win=200, avg_loss=100),
                                                           dict(length=200, hit_rate=0.5, avg_win=100, avg_loss=100),
                                                           dict(length=500, hit_rate=0.2, avg_win=100, avg_loss=200)])}
        extras["other_configs"] = {}
        for name, kw in kws.items():
            pm = make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, SIMS_PER_GPU_STEP // 2, NUM_SHUFFLES, **kw)
            for sd in (1, 2):
                eng.simulate(pm, seed=sd, options=eng.options(chunk_paths=CHUNK_PATHS))
            t = eng.timing()
            for sd in (3, 4):   # first call loads the kernel; the second is timed
                eng.simulate(pm, seed=sd, options=eng.options(chunk_paths=CHUNK_PATHS, int_fast_path=False))
            t64 = eng.timing()
            extras["other_configs"][name] = {"value": t["trade_steps"] / (t["total_ms"] * 1e-3),
                                             "value_fp64_evaluator": t64["trade_steps"] / (t64["total_ms"] * 1e-3), "unit": UNIT}
        # (a2b) worst case for generation: ONE shuffle per simulation, so the base sequence (1000 53-bit uniforms) is not amortised
        extras["gen_worst_case"] = {}
        for name, kw in (("dynamic_fp64", {}), ("markov2_int32", kws["markov2_c3"])):
            pm = make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, 2097140, 1, **kw)
            for sd in (1, 2):
                eng.simulate(pm, seed=sd, options=eng.options(chunk_paths=CHUNK_PATHS, serial_chunks=True))
            t = eng.timing()
            extras["gen_worst_case"][name] = {"value": t["trade_steps"] / (t["total_ms"] * 1e-3), "unit": UNIT, "num_mc_shuffles": 1,
                                              "gen_base_share_of_step": t["gen_ms"] / max(t["total_ms"], 1e-9)}
        # (a3) replay mode end to end: caller sequences in (pageable) host memory -> C ABI -> per-strategy aggregates
        import numpy as np
        n_rep = 200000
        seqs = np.ascontiguousarray(eng.sequences(make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, n_rep // NUM_SHUFFLES, NUM_SHUFFLES), seed=5))
        eng.replay(seqs, per_path=False)
        t0 = time.perf_counter()
        eng.replay(seqs, per_path=False)
        wall = time.perf_counter() - t0
        t = eng.timing()
        extras["replay"] = {"e2e_value": t["trade_steps"] / wall, "kernels_value": t["trade_steps"] / (t["paths_ms"] * 1e-3), "unit": UNIT,
                            "h2d_bytes": int(t["h2d_bytes"]), "d2h_bytes": int(t["d2h_bytes"]), "paths": n_rep,
                            "note": "replay of caller-supplied fp64 sequences: 8 B per (path, trade) cross PCIe from pageable host memory, staged by "
                                    "worker threads through pinned slices (csrc/stager.hpp)"}
        del seqs
        # (b) equity-curve export: HBM-bound, 8 B per trade-step
        n_curve = 1 << 17
        dev = eng.device_alloc(20 * ((NUM_TRADES + 1) // 2) * n_curve * 16)      # 21 GB, written three times; the last run counts
        for sd in (3, 4, 5):
            eng.export_curves(small, sd, 0, n_curve, to_host=False, dev=dev)
        t = eng.timing()
        eng.device_free(dev)
        gbs = t["trade_steps"] * 8 / (t["paths_ms"] * 1e-3) / 1e9
        extras["curve_export"] = {"trade_steps_per_s": t["trade_steps"] / (t["paths_ms"] * 1e-3), "achieved_gbs": gbs,
                                  "bytes_per_trade_step": 8, "paths": n_curve}

    # ---- lines every rank takes part in: the fixed-size problems (BASELINE.json config 5 is a STRONG-scaling statement) ----
    import hashlib

    def sharded_rows(pm, seed, chunk=0):
        """run_all_strategies on all N GPUs of this job through the public entry points (every rank gets the table)."""
        if world == 1:
            return eng.simulate(pm, seed=seed, options=eng.options(chunk_paths=chunk))["summary"], eng.timing()["total_ms"]
        ch = mdist.resolve_chunk(pm.num_simulations * pm.num_mc_shuffles, chunk, world)
        _, rows = mdist.simulate_sharded(pm, seed, lambda lo_, hi_: eng.simulate_chunks(pm, seed, lo_, hi_, options=eng.options(chunk_paths=ch)),
                                         rank, world, chunk_paths=ch, device=device)
        return rows, eng.timing()["total_ms"]

    def timed_job(pm, seed, reps, vary_seed=True):
        best = None
        for r in range(reps):
            barrier()
            t0 = time.perf_counter()
            rows, dev_ms_local = sharded_rows(pm, seed + (r if vary_seed else 0))
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            tt = torch.tensor([wall, dev_ms_local * 1e-3], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if best is None or float(tt[0]) < best[0]:
                best = (float(tt[0]), float(tt[1]), rows)
        return best

    if not args.no_extras:
        # (c) strong scaling: ONE fixed problem (67,109,120 paths) split over however many GPUs there are
        pm = make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, STRONG_SIMS, NUM_SHUFFLES)
        wall, dev_s, rows = timed_job(pm, 31000, 3)
        tsteps = STRONG_SIMS * NUM_SHUFFLES * NUM_TRADES * 20
        extras["strong_scaling"] = {"paths_total": STRONG_SIMS * NUM_SHUFFLES, "n_gpus": n_gpus, "ms_per_step": wall * 1e3, "device_ms_per_step_max_over_ranks": dev_s * 1e3,
                                    "value": tsteps / wall, "unit": UNIT, "scaling": "strong",
                                    "note": "fixed problem; wall time of the whole sharded call (best of 3), max over ranks"}
        # (d) BASELINE.json config 5 at its stated size: 1e9 paths x 1000 trades x 20 strategies
        pm = make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, C5_FULL_SIMS, NUM_SHUFFLES)
        wall, dev_s, rows = timed_job(pm, 2026, 2, vary_seed=False)   # best of 2: a single 1e9-path call is exposed to host hiccups
        extras["c5_full"] = {"paths": C5_FULL_SIMS * NUM_SHUFFLES, "trade_steps": C5_FULL_SIMS * NUM_SHUFFLES * NUM_TRADES * 20, "n_gpus": n_gpus, "wall_s": wall,
                             "value": C5_FULL_SIMS * NUM_SHUFFLES * NUM_TRADES * 20 / wall, "unit": UNIT, "seed": 2026,
                             "avg_profit_strategy1": float(rows[0][0]), "avg_drawdown_strategy1": float(rows[0][1]), "analytic_avg_profit_strategy1": 103063.2,
                             "summary_sha256": hashlib.sha256(np_bytes(rows)).hexdigest()}
    # (e) partition invariance, driver-visible: the summary of one fixed 2^24-path run must hash the same for every N
    pm = make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, SHA_SIMS, SHA_SHUFFLES)
    rows, _ = sharded_rows(pm, 4711)
    partition_sha = hashlib.sha256(np_bytes(rows)).hexdigest()

    if rank != 0:
        if world > 1:
            dist.barrier(group=host_group)   # rank 0 may still be running its single-process measurements (CPU-side wait)
            dist.barrier()
            dist.destroy_process_group()
        return

    # (f) the same fixed run through the UNCHANGED API from ONE process driving all N GPUs inside libmcts (no torchrun
    #     needed by a user; the other ranks of this job idle on a CPU-side barrier meanwhile)
    inlib = None
    try:
        from mc_trade_sizing_b200 import EngineGroup
        with EngineGroup(list(range(n_gpus))) as grp:
            grp.simulate(pm, seed=1)
            t0 = time.perf_counter()
            out = grp.simulate(pm, seed=4711)
            wall = time.perf_counter() - t0
            inlib = {"devices": grp.devices, "summary_sha256": hashlib.sha256(np_bytes(out["summary"])).hexdigest(), "wall_ms": wall * 1e3,
                     "value": SHA_SIMS * SHA_SHUFFLES * NUM_TRADES * 20 / wall, "unit": UNIT,
                     "api": "mcts_group_simulate behind run_all_strategies (MCTS_DEVICES): one process, one host thread per device"}
            if not args.no_extras and n_gpus > 1:
                pm_s = make_params(HIT_RATE, AVG_WIN, AVG_LOSS, NUM_TRADES, STRONG_SIMS, NUM_SHUFFLES)
                grp.simulate(pm_s, seed=2)
                t0 = time.perf_counter()
                grp.simulate(pm_s, seed=3)
                inlib["strong_scaling_ms"] = (time.perf_counter() - t0) * 1e3
    except Exception as e:
        inlib = {"error": repr(e)}

    peaks_file = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks_file = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
    if "curve_export" in extras:
        extras["curve_export"]["roofline"] = {"bound": "hbm", "achieved": extras["curve_export"]["achieved_gbs"], "peak": hbm_peak,
                                              "unit": "GB/s", "frac": extras["curve_export"]["achieved_gbs"] / hbm_peak,
                                              "peak_source": "MEASURED_PEAKS.json" if peaks_file else "fallback 6.65 TB/s"}

    kernel_rate = kstat["trade_steps"] / (kstat["paths_ms"] * 1e-3)  # this rank's dominant kernel, launches serialised
    achieved_tflops = FP64_OPS_PER_TRADE_STEP * kernel_rate / 1e12
    peak_tflops = peaks["fp64_inst_per_s"] / 1e12
    roofline = {
        "bound": "fp64-issue", "kernel": "k_sim_paths", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": achieved_tflops / peak_tflops,
        # dram__bytes_read.sum + dram__bytes_write.sum of one k_sim_paths launch of this workload (2,097,160 paths), from the
        # committed `ncu --set full` capture profiles/r01_ncu_bench_sim_paths.txt; algorithmic bytes per launch = paths x 320 B
        "traffic": TRAFFIC_PER_LAUNCH, "traffic_unit": "B per launch", "traffic_source": TRAFFIC_SOURCE, "algorithmic_bytes_per_launch": 2097160 * 320,
        "peak_source": "in-process microbenchmark k_peak_dadd (mcts_measure_peaks): independent DADD chains over the whole chip, measured by "
                       "this run; MEASURED_PEAKS.json has no fp64 figure",
        "bound_note": "neither HBM (0.016 B per trade-step) nor tensor cores (no contraction) bound this kernel; SURVEY 8(d) "
                      "prescribes the SM issue / FP64-pipe roofline used here; the HBM roofline of curve-export mode is under curve_export",
        "definition": "achieved = 6 algorithmic fp64 ops per trade-step (SURVEY 8d) x trade-steps per launch / mean launch "
                      "duration (CUDA events on the launching stream, %d launches serialised on one stream right after the timed region); "
                      "peak = fp64 DADD instruction rate measured in this process by mcts_measure_peaks (MEASURED_PEAKS.json has "
                      "no fp64 figure), FMA not counted double" % kstat["paths_launches"],
        "kernel_ms_per_launch": kstat["paths_ms"] / max(kstat["paths_launches"], 1),
        "kernel_share_of_step": kstat["paths_ms"] / max(kstat["total_ms"], 1e-9),
        "int_alu_inst_per_s": peaks["int_inst_per_s"], "fp64_inst_per_s": peaks["fp64_inst_per_s"],
        "sm_clock_mhz_microbench": peaks["sm_clock_mhz"],
        # register-only replica of the SELECT-based trade step (3 DADD + 2 DSETP + 5 FSEL per strategy-step, no generator, no
        # shared memory, no stores) at the kernel's occupancy, measured in this process (mcts_measure_step_ceiling).  It
        # saturates the ALU pipe; the shipped kernel moves the peak / drawdown selects to the FMA pipe as predicated IMADs
        # (DESIGN.md section 5), so it is no longer bounded by this figure -- kept as a reference point
        "mix_ceiling": {"strategy_steps_per_s": peaks["step_ceiling_per_s"], "kernel_frac": kernel_rate / peaks["step_ceiling_per_s"],
                        "note": "select-based replica (ALU-pipe bound); the kernel uses FMA-pipe predicated moves instead"},
        # register-only loop with the shipped step's instruction mix (5 DADD + 4 IMAD + 4 LOP3 per strategy-step, independent chains) at
        # the kernel's occupancy, measured in this process (mcts_measure_mix_ceiling): what the schedulers can issue for this mix
        "issue_mix_ceiling": {"strategy_steps_per_s": peaks["mix_ceiling_per_s"], "ipc_per_scheduler": peaks["mix_ceiling_ipc"],
                              "kernel_frac": kernel_rate / peaks["mix_ceiling_per_s"],
                              "note": "13 instructions per step in the replica (ptxas schedules it D F D A ...: an fp64 instruction in every other "
                                      "slot), ~13.5 in the kernel incl. generation; the same mix in ptxas' clustered schedule issues 0.845 "
                                      "(profiles/r02_pipe_mix_bench.txt); the kernel's own IPC is 0.78 (profiles/r02_ncu_bench_sim_paths.txt)"},
        "hbm_note": "aggregate mode moves 0.016 B per trade-step (per-path results only); HBM bounds only curve export",
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(n_gpus),
        "wall_ms_per_step": wall_ms / args.steps,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(t_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(t_e2e["d2h_bytes"]),
                "api": "mc_trade_sizing_b200.run_all_strategies (ctypes -> libmcts C ABI, host buffers)" if world == 1 else
                       "mc_trade_sizing_b200.distributed.run_all_strategies_sharded",
                "note": "inputs are the scalar arguments of run_all_strategies (kernel parameter blocks), outputs the per-chunk "
                        "aggregate records; trade outcomes are generated on the device"},
        "gpu_launches": int(stat["launches"]),
        "clocks": clocks, "roofline": roofline,
        "result_check": {"avg_profit_strategy1": float(last[1][0][0]), "expected": 103063.2},
        "partition_sha": partition_sha,
        "partition_sha_of": "sha256 of the [20,9] summary of the fixed run %d sims x %d shuffles x 1000 trades, seed 4711, sharded over this job's %d GPU(s): "
                            "must be identical for every N" % (SHA_SIMS, SHA_SHUFFLES, n_gpus),
        "inlib_multi_gpu": inlib,
    }
    line.update(extras)
    if not args.no_cpu_baseline and n_gpus == 1:
        line["cpu_baseline"] = {k: v for k, v in cpu_reference_rate().items() if k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["python_reference"] = python_reference_rates()
    if not args.no_extras and n_gpus == 1:
        line["dps_py_unchanged"] = dps_overlay_walls()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(group=host_group)
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
