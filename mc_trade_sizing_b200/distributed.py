"""Multi-GPU sharding of one run_all_strategies call: one process per GPU (torch.distributed).

Paths are independent, so the path range is cut into nodes of NumPy's pairwise-summation tree
(`Engine.plan`), contiguous runs of nodes go to the ranks, and ONE collective at the end exchanges
the ~1 KB per-node records (`mcts_chunk_result`); every rank then folds them in tree order on the
host (`mcts_combine`), which makes the result bit-identical for any number of GPUs.  There is no
data-path collective: Philox draws are keyed by (simulation, shuffle, trade).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .engine import Engine, agg_to_rows

REC_BYTES = C.sizeof(_ffi.ChunkResult)


def shard_bounds(n_chunks: int, world: int):
    """Contiguous chunk ranges [lo, hi) per rank, balanced to within one chunk (by chunk COUNT; the sharded
    entry points below balance by path count through Engine.plan_shards)."""
    return [(n_chunks * r // world, n_chunks * (r + 1) // world) for r in range(world)]


def resolve_chunk(P: int, chunk_paths: int, world: int) -> int:
    """chunk_paths == 0 with several ranks: a chunk size that gives every rank at least four chunks.  The
    single-device default (1 << 22) would leave any run of <= 4.19 M paths on ONE rank and an 8.39 M-path run on
    four of eight.  Results are bit-identical for every chunk size; plan, evaluation and combine must merely agree
    on it, which is why the resolved value is returned and used for all three."""
    return int(chunk_paths) if chunk_paths and chunk_paths > 0 else (Engine.auto_chunk(P, world) if world > 1 else 0)


def records_to_bytes(recs, n: int) -> np.ndarray:
    buf = np.zeros(n * REC_BYTES, dtype=np.uint8)
    if n:
        C.memmove(buf.ctypes.data, C.addressof(recs), n * REC_BYTES)
    return buf


def bytes_to_records(buf: np.ndarray, n: int):
    recs = (_ffi.ChunkResult * max(n, 1))()
    if n:
        C.memmove(C.addressof(recs), np.ascontiguousarray(buf).ctypes.data, n * REC_BYTES)
    return recs


_PINNED = {}


def _pinned(nbytes: int, tag: str):
    """Reusable pinned host staging buffers (the exchange is a few KB; allocation would dominate it)."""
    import torch
    buf = _PINNED.get(tag)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8).pin_memory()
        _PINNED[tag] = buf
    return buf[:nbytes]


def gather_records(local_bytes: np.ndarray, counts, rank: int, world: int, device=None, group=None) -> np.ndarray:
    """all_gather of the per-rank record blocks (padded to the largest block); returns the
    concatenation in rank order.  `device` None -> CPU tensors (gloo), else CUDA tensors (NCCL): pinned staging
    buffers, asynchronous copies on the current stream and ONE synchronisation, after the gathered block is back."""
    import torch
    import torch.distributed as dist
    width = max(max(counts) * REC_BYTES, 1)
    if device is None:
        send = torch.zeros(width, dtype=torch.uint8)
        send[: local_bytes.size] = torch.from_numpy(local_bytes)
        recv = torch.empty(world * width, dtype=torch.uint8)
        dist.all_gather_into_tensor(recv, send, group=group)
        host = recv.numpy().reshape(world, width)
    else:
        stage = _pinned(width, "send")
        stage.zero_()
        stage[: local_bytes.size] = torch.from_numpy(local_bytes)
        send = stage.to(device, non_blocking=True)
        recv = torch.empty(world * width, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(recv, send, group=group)
        back = _pinned(world * width, "recv")
        back.copy_(recv, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        host = back.numpy().reshape(world, width)
    return np.concatenate([host[r, : counts[r] * REC_BYTES] for r in range(world)])


def simulate_sharded(params, seed: int, compute_chunks, rank: int, world: int, chunk_paths: int = 0, device=None, group=None):
    """compute_chunks(lo, hi) -> (ChunkResult array, n) evaluates plan nodes [lo, hi) locally
    (Engine.simulate_chunks on a GPU).  Returns (agg, summary rows [20,9]) on every rank."""
    P = params.num_simulations * params.num_mc_shuffles
    if P <= 0:
        raise ValueError("zero paths")
    if world > 1 and not (chunk_paths and chunk_paths > 0):
        raise ValueError("simulate_sharded needs the resolved chunk size (resolve_chunk) so that compute_chunks uses the same plan")
    first, _ = Engine.plan(P, chunk_paths)
    n_chunks = len(first)
    bounds = Engine.plan_shards(P, chunk_paths, world)
    lo, hi = bounds[rank]
    recs, n = compute_chunks(lo, hi)
    counts = [b - a for a, b in bounds]
    if world > 1:
        allb = gather_records(records_to_bytes(recs, n), counts, rank, world, device=device, group=group)
    else:
        allb = records_to_bytes(recs, n)
    agg = Engine.combine(P, chunk_paths, bytes_to_records(allb, n_chunks), n_chunks)
    return agg, agg_to_rows(agg, params.num_trades)


def run_all_strategies_sharded(engine: Engine, params, seed: int, rank: int, world: int, chunk_paths: int = 0, options=None,
                               device=None, group=None):
    """The N-GPU form of run_all_strategies: every rank returns the same ranked 20-row table."""
    from .trading_models import summary_to_rows
    P = params.num_simulations * params.num_mc_shuffles
    chunk_paths = resolve_chunk(P, chunk_paths, world)
    opt = options if options is not None else engine.options(chunk_paths=chunk_paths)
    opt.chunk_paths = chunk_paths

    def compute(lo, hi):
        return engine.simulate_chunks(params, seed, lo, hi, options=opt)

    _, rows = simulate_sharded(params, seed, compute, rank, world, chunk_paths=chunk_paths, device=device, group=group)
    return summary_to_rows(rows)
