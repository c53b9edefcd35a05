"""World-size-2 test of the sharding logic on CPU (gloo): plan -> contiguous tree-node ranges per
rank -> one all_gather of the ~1 KB records -> tree combine.  The per-chunk compute is the CPU
harness of the device code (test only); on GPUs it is Engine.simulate_chunks + NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _chunk_records(prm, seed, first, count, lo, hi):
    import harness
    from mc_trade_sizing_b200 import _ffi
    n = hi - lo
    recs = (_ffi.ChunkResult * max(n, 1))()
    for c in range(lo, hi):
        out = harness.simulate(prm, seed, first_path=int(first[c]), n=int(count[c]))
        r = recs[c - lo]
        r.first_path, r.n_paths = int(first[c]), int(count[c])
        for s in range(20):
            r.sum_profit[s] = harness.pairwise(out["profit"][s])
            r.sum_dd[s] = harness.pairwise(out["dd"][s])
            r.min_profit[s], r.max_profit[s] = out["profit"][s].min(), out["profit"][s].max()
            r.min_dd[s], r.max_dd[s] = out["dd"][s].min(), out["dd"][s].max()
    return recs, n


def _worker(rank, world, port, chunk, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mc_trade_sizing_b200 import Engine, make_params
    from mc_trade_sizing_b200 import distributed as mdist
    prm = make_params(0.82, 186.0, 219.0, 64, 130, 7)
    first, count = Engine.plan(910, chunk)
    agg, rows = mdist.simulate_sharded(prm, 77, lambda lo, hi: _chunk_records(prm, 77, first, count, lo, hi), rank, world,
                                       chunk_paths=chunk)
    q.put((rank, rows.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,chunk", [(2, 128), (2, 512), (3, 128)])
def test_sharded_run_equals_single_process(world, chunk):
    import harness
    from mc_trade_sizing_b200 import make_params
    from oracle import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prm = make_params(0.82, 186.0, 219.0, 64, 130, 7)
    full = harness.simulate(prm, 77)
    ref = oracle.summarise(full["profit"], full["dd"], 64)
    for r in range(world):
        rows = np.frombuffer(got[r], dtype=np.float64).reshape(20, 9)
        assert (rows.view(np.uint64) == ref.view(np.uint64)).all(), "rank %d differs from the single-process result" % r


def test_shard_bounds_cover_all_chunks():
    from mc_trade_sizing_b200.distributed import shard_bounds
    for n in (1, 2, 7, 8, 256):
        for world in (1, 2, 3, 4, 8):
            b = shard_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_auto_chunk_gives_every_rank_work():
    """with the single-device default chunk (1 << 22) a run of <= 4.19 M paths is ONE chunk and an
    8.39 M-path run is four, so most of eight ranks idled.  resolve_chunk + plan_shards must hand every rank a
    contiguous, non-empty, balanced run whenever there are enough paths, and cover the plan exactly."""
    from mc_trade_sizing_b200 import Engine
    from mc_trade_sizing_b200.distributed import resolve_chunk
    for P in (40_000, 1_000_000, 4_194_304, 8_388_640, 67_109_120, 10 ** 9):
        for world in (2, 3, 4, 8):
            ch = resolve_chunk(P, 0, world)
            first, count = Engine.plan(P, ch)
            b = Engine.plan_shards(P, ch, world)
            assert b[0][0] == 0 and b[-1][1] == len(first) and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            per = [int(count[lo:hi].sum()) for lo, hi in b]
            assert sum(per) == P
            if P >= 16384 * 4 * world:
                assert min(per) > 0, (P, world, per)
                assert max(per) * world <= 1.3 * P, (P, world, per)
    assert resolve_chunk(10 ** 6, 0, 1) == 0 and resolve_chunk(10 ** 6, 4096, 8) == 4096
