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


def test_shard_chunks_cover_every_point_once_at_every_world_size():
    for total, chunk in ((100_000_000, 500_000), (1_000_003, 4096), (5, 7), (0, 10)):
        for world in (1, 2, 4, 8):
            seen = []
            per_rank = []
            for r in range(world):
                mine = list(sdist.shard_chunks(total, chunk, r, world))
                per_rank.append(sum(n for _, _, n in mine))
                seen += mine
            seen.sort()
            assert [c for c, _, _ in seen] == list(range((total + chunk - 1) // chunk))  # every chunk exactly once
            assert all(first == c * chunk and 0 < n <= chunk for c, first, n in seen)
            assert sum(per_rank) == total
            if total == 100_000_000:  # the sweep of bench.py: the same number of points on every rank at 1, 2, 4 and 8 GPUs
                assert len(set(per_rank)) == 1


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


def test_shard_columns_is_aligned_partition():
    for d_out in (1, 7, 8, 100, 10_000, 10_001):
        for world in (1, 2, 3, 8):
            blocks = [sdist.shard_columns(d_out, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == d_out
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert all(lo % 8 == 0 or lo == d_out for lo, _ in blocks)


def _column_worker(rank, world, port, out):
    import torch
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator
    from test_plan import DENSE, Plan

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smolyax_b200 import workloads

        w = workloads.Workload("wide", "leja" if rank >= 0 else "gh", 6, 20, 120, 0)
        d_in, d_out = w.d_in, w.d_out
        full, _ = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=d_out)._assemble_compact(w.target(), {})
        local = sdist.scatter_columns(full if rank == 0 else None, d_out, src=0)
        lo, hi = sdist.shard_columns(d_out, rank, world)
        assert local["values"].shape == (full["values"].shape[0], hi - lo)
        x = w.points(40, seed=3)
        y_full = Plan(full, d_in, d_out, DENSE).dense(x)
        y_loc = Plan(local, d_in, hi - lo, DENSE).dense(x)
        same = np.array_equal(y_loc, y_full[:, lo:hi])
        again = np.array_equal(Plan(sdist.column_slice(full, lo, hi), d_in, hi - lo, DENSE).dense(x), y_loc)
        y_all = sdist.gather_columns(torch.from_numpy(y_loc), d_out).numpy()
        out[rank] = (same, again, np.array_equal(y_all, y_full), lo, hi)
    finally:
        dist.destroy_process_group()


def test_column_sharding_gloo():
    """d_out sharded over two ranks: each rank receives only its slice of the value table, evaluates its outputs, and
    the gathered result equals the unsharded one bit for bit (plan level, CPU)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_column_worker, args=(2, port, out), nprocs=2, join=True)
    for r in (0, 1):
        assert out[r][0] and out[r][1] and out[r][2], out[r]
    assert (out[0][3], out[0][4], out[1][3], out[1][4]) == (0, 16, 16, 20)
