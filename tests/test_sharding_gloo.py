"""World-size-2 CPU test of the N>1 host path: band assignment + merge of disjoint rows by a SUM reduce
(the collective bench.py uses over NCCL), run over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fractalshark_b200.sharding import merge_shards, rows_of_shard


def test_bands_partition_the_frame():
    for h in (1, 4, 37, 2160):
        for n in (1, 2, 3, 8):
            seen = np.concatenate([rows_of_shard(h, n, r) for r in range(n)])
            assert sorted(seen.tolist()) == list(range(h))
    assert rows_of_shard(16, 2, 1).tolist() == [4, 5, 6, 7, 12, 13, 14, 15]


def _worker(rank, world, port, h, w, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1234)
    full = rng.integers(1, 1 << 20, size=(h, w), dtype=np.int64)   # same "frame" on every rank
    mine = np.zeros_like(full)
    rows = rows_of_shard(h, world, rank)
    mine[rows] = full[rows]                                        # a rank only renders its own bands
    total = torch.tensor([int(mine.sum())])                        # this rank's checksum
    t = torch.from_numpy(mine)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)                    # disjoint rows: SUM == gather
    dist.all_reduce(total)
    if rank == 0:
        out_q.put((np.array_equal(t.numpy(), full), int(total.item()) == int(full.sum())))
    dist.destroy_process_group()


def test_two_rank_reduce_is_a_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, 50, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok == (True, True)


def test_merge_shards_helper():
    h, w = 21, 16
    full = np.arange(h * w).reshape(h, w)
    bufs = []
    for r in range(3):
        b = np.zeros_like(full)
        rows = rows_of_shard(h, 3, r)
        b[rows] = full[rows]
        bufs.append(b)
    assert np.array_equal(merge_shards(bufs, h), full)


def _replicate_worker(rank, world, port, out_q):
    """Rank 0 produces orbit + LA table + coordinates, packs them, broadcasts the blobs (the way bench.py does over
    NCCL); rank 1 unpacks without computing anything and reports the CRC of what it would upload."""
    import zlib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fractalshark_b200 import Numeric
    from fractalshark_b200.host_inputs import LaTable, Orbit, ReplicatedInputs, View
    from fractalshark_b200.views import VIEW5

    def crc_of(coords, orbit, la):
        crc = 0
        for k in sorted(coords):
            crc = zlib.crc32(coords[k], crc)
        d, l = orbit.descriptor(), la.descriptor()
        import ctypes as C
        for ptr, n in ((d.elements, orbit.count * orbit.elem_bytes), (l.las, la.num_las * la.las_elem_bytes),
                       (l.stages, la.num_stages * 2 * la.iter_bytes), (l.at, la.at_bytes)):
            crc = zlib.crc32(bytes((C.c_ubyte * n).from_address(ptr)), crc)
        return crc, (d.compressed_count, d.uncompressed_count, d.period_maybe_zero, l.la_stage_count, l.use_at, l.is_valid)

    box = [None]
    if rank == 0:
        v = View(VIEW5.min_x, VIEW5.min_y, VIEW5.max_x, VIEW5.max_y, 64, 36)
        coords = v.coords(Numeric.HDR32)
        orbit = Orbit(v, Numeric.HDR32, VIEW5.num_iterations, True)
        la = LaTable(orbit, 4)
        meta, blobs = ReplicatedInputs.pack(coords, orbit, la, VIEW5.num_iterations)
        box = [meta]
        out_q.put(("src", crc_of(coords, orbit, la)))
    dist.broadcast_object_list(box, src=0)
    meta = box[0]
    got = []
    for i, size in enumerate(meta["sizes"]):
        t = torch.from_numpy(blobs[i]) if rank == 0 else torch.empty(size, dtype=torch.uint8)
        if size:
            dist.broadcast(t, src=0)
        got.append(t.numpy())
    if rank == 1:
        coords, orbit, la, n = ReplicatedInputs.unpack(meta, got)
        out_q.put(("dst", crc_of(coords, orbit, la), n))
    dist.destroy_process_group()


def test_inputs_replicate_by_broadcast_without_recomputation():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_replicate_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict()
    for _ in range(2):
        item = q.get(timeout=180)
        res[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
    assert res["src"][0] == res["dst"][0]
    assert res["dst"][1] == 4718592
