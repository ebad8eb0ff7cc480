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
