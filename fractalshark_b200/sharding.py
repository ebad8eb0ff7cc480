"""Multi-GPU sharding of the frame: 4-row tile bands are dealt round-robin to ranks.

The same mapping is compiled into the kernels (`fs_set_shard`, fs_lav2.cuh / fs_direct.cuh); this module
is the host-side statement of it used to merge / validate per-rank buffers.  No data-path collective is
needed inside the render: pixels are independent and the inputs (orbit, LA table) are read-only replicas.
"""
from __future__ import annotations

import numpy as np

BAND_ROWS = 4


def rows_of_shard(height: int, shard_count: int, shard_index: int) -> np.ndarray:
    """Row indices rendered by `shard_index`."""
    rows = np.arange(height)
    return rows[(rows // BAND_ROWS) % shard_count == shard_index]


def merge_shards(buffers: list[np.ndarray], height: int) -> np.ndarray:
    """Assemble the full iteration buffer from per-rank buffers (each holds only its own bands)."""
    out = np.zeros_like(buffers[0])
    n = len(buffers)
    for r, b in enumerate(buffers):
        rows = rows_of_shard(height, n, r)
        out[rows] = b[rows]
    return out
