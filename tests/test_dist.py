"""Multi-process plumbing of the evaluation path on CPU (gloo, world size 2): row sharding and the one-off broadcast
of the tables from rank 0.  The per-call path has no collective, so this is all there is to test without GPUs."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from smolyax_b200 import dist as sdist
from helpers import golden_layout, load


def test_shard_rows_partitions_exactly():
    for n in (0, 1, 7, 32, 1_000_003):
        for world in (1, 2, 3, 8):
            blocks = [sdist.shard_rows(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layout = golden_layout(load("medium_00")) if rank == 0 else None
        got = sdist.broadcast_layout(layout, src=0)
        ref = golden_layout(load("medium_00"))
        ok = set(got) == set(ref) and all(np.array_equal(got[k], ref[k]) and got[k].dtype == ref[k].dtype for k in ref)
        lo, hi = sdist.shard_rows(101, rank, world)
        slowest = sdist.max_over_ranks(float(rank + 1))
        out[rank] = (ok, lo, hi, slowest)
    finally:
        dist.destroy_process_group()


def test_broadcast_layout_and_max_over_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0][0] and out[1][0]
    assert (out[0][1], out[0][2], out[1][1], out[1][2]) == (0, 51, 51, 101)
    assert out[0][3] == out[1][3] == 2.0
