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


def band_copy_plan(padded_rows: int, shard_count: int, shard_index: int) -> tuple[int, int, int]:
    """``(first_row, row_stride, n_bands)`` of the one strided 2-D copy ``fs_render_current_shard`` issues: bands of
    ``BAND_ROWS`` rows starting at ``first_row`` and repeating every ``row_stride`` rows (``padded_rows`` is the
    buffer height, a multiple of 8)."""
    bands = padded_rows // BAND_ROWS
    owned = (bands - shard_index + shard_count - 1) // shard_count
    return shard_index * BAND_ROWS, shard_count * BAND_ROWS, max(owned, 0)


class SharedFrame:
    """One whole-frame iteration buffer in POSIX shared memory that every rank of a node writes its own bands into
    (result sink or ``RenderCurrentShard``): the frame is assembled on the host with no collective.  Rank 0 creates
    and finally unlinks the segment; the other ranks attach after a barrier of the caller's choosing."""

    def __init__(self, name: str, shape: tuple[int, int], dtype=np.uint32, create: bool = False):
        from multiprocessing import shared_memory
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.owner = create
        if create:
            try:
                shared_memory.SharedMemory(name=name).unlink()  # a stale segment of a crashed run
            except FileNotFoundError:
                pass
            self._shm = shared_memory.SharedMemory(name=name, create=True, size=nbytes)
        else:
            self._shm = shared_memory.SharedMemory(name=name)
            # the creator owns the segment: keep this process's resource tracker from unlinking it again at exit
            from multiprocessing import resource_tracker
            resource_tracker.unregister(self._shm._name, "shared_memory")
        self.array = np.ndarray(shape, dtype=dtype, buffer=self._shm.buf)
        if create:
            self.array[:] = 0

    def close(self) -> None:
        self.array = None
        self._shm.close()
        if self.owner:
            self._shm.unlink()
