"""World-size-2 CPU test of the N>1 host path: band assignment + merge of disjoint rows by a SUM reduce
(the collective bench.py uses over NCCL), run over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fractalshark_b200.sharding import SharedFrame, band_copy_plan, merge_shards, rows_of_shard


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


def _frame_worker(rank, world, port, h, w, out_q):
    """The host side of the multi-GPU result path: every rank writes only its bands, with the strided-copy geometry of
    fs_render_current_shard, into ONE shared-memory frame; rank 0 reads the whole picture.  No data-path collective."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hp, wp = (h + 7) // 8 * 8, (w + 15) // 16 * 16                  # padded like the library's buffers
    rng = np.random.default_rng(99)
    full = rng.integers(1, 1 << 30, size=(hp, wp), dtype=np.uint32)  # the frame every rank would render
    name = "fsb200_test_frame_%d" % port
    shm = SharedFrame(name, (hp, wp), np.uint32, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        shm = SharedFrame(name, (hp, wp), np.uint32)
    first, stride, n_bands = band_copy_plan(hp, world, rank)
    for b in range(n_bands):                                        # what the one cudaMemcpy2DAsync does
        r0 = first + b * stride
        shm.array[r0:r0 + 4] = full[r0:r0 + 4]
    dist.barrier()
    if rank == 0:
        owned = [set(range(f + b * s, f + b * s + 4)) for (f, s, n) in (band_copy_plan(hp, world, k) for k in range(world))
                 for b in range(n)]
        disjoint_cover = sum(len(o) for o in owned) == hp and set().union(*owned) == set(range(hp))
        rows_agree = all(np.array_equal(np.array(sorted(set().union(*[set(range(f + b * s, f + b * s + 4)) for b in range(n)]) & set(range(h)))),
                                        rows_of_shard(h, world, k))
                         for k, (f, s, n) in ((k, band_copy_plan(hp, world, k)) for k in range(world)))
        out_q.put((bool(np.array_equal(shm.array, full)), disjoint_cover, rows_agree))
    dist.barrier()
    shm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("h,w", [(37, 50), (64, 48)])
def test_two_ranks_assemble_one_shared_host_frame(h, w):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_frame_worker, args=(r, 2, port, h, w, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok == (True, True, True)
    assert all(p.exitcode == 0 for p in procs)


def test_band_copy_plan_matches_row_ownership():
    for hp in (8, 40, 2160):
        for n in (1, 2, 3, 8):
            rows = []
            for k in range(n):
                first, stride, bands = band_copy_plan(hp, n, k)
                mine = [r for b in range(bands) for r in range(first + b * stride, first + b * stride + 4)]
                assert mine == rows_of_shard(hp, n, k).tolist()
                rows += mine
            assert sorted(rows) == list(range(hp))


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
