"""Equity-curve export (an extension -- the reference only has the temporary
`cumulative_equity` of src/trading_models.py:167 and never writes it).

Device layout written by `mcts_export_curves`: [20 strategies][ceil(L/2)][n paths][2 trades] fp64 --
two consecutive trades per 128-bit store, paths contiguous, so a warp's stores are 512 contiguous
bytes.  `curves_array` returns it reshaped to [20, n, L]; `write_parquet` stores a long table next to
what `src/output_handler.save_parquet` writes for the result table (reference src/output_handler.py:84-93).
"""
from __future__ import annotations

import numpy as np

from .engine import Engine
from .trading_models import DESCRIPTIONS


def curves_array(engine: Engine, params, seed: int, first_path: int, n: int) -> np.ndarray:
    """Equity curves [20, n, L] of paths [first_path, first_path + n)."""
    raw = engine.export_curves(params, seed, first_path, n)                 # [20, half, n, 2]
    L = engine.sequences(params, seed, first_path, 1).shape[1]
    return np.ascontiguousarray(raw.transpose(0, 2, 1, 3).reshape(20, n, -1)[:, :, :L])


def write_parquet(engine: Engine, params, seed: int, first_path: int, n: int, filename: str, strategies=None) -> int:
    """Write a long table (path, strategy_id, strategy, trade, equity) for the selected strategy ids
    (default all 20).  Returns the number of rows written."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    curves = curves_array(engine, params, seed, first_path, n)
    ids = list(range(1, 21)) if strategies is None else [int(s) for s in strategies]
    L = curves.shape[2]
    writer = None
    rows = 0
    try:
        for sid in ids:
            tbl = pa.table({
                "path": np.repeat(np.arange(first_path, first_path + n, dtype=np.int64), L),
                "strategy_id": np.full(n * L, sid, dtype=np.int16),
                "strategy": pa.array([DESCRIPTIONS[sid]] * (n * L)).dictionary_encode(),
                "trade": np.tile(np.arange(L, dtype=np.int32), n),
                "equity": curves[sid - 1].reshape(-1),
            })
            if writer is None:
                writer = pq.ParquetWriter(filename, tbl.schema)
            writer.write_table(tbl)
            rows += tbl.num_rows
    finally:
        if writer is not None:
            writer.close()
    return rows
