"""Multi-GPU plumbing for the evaluation path: one process per GPU, points partitioned, tables replicated.

The path shards by evaluation point with no per-call collective: rank r evaluates the contiguous
row block ``shard_rows(N, r, world)`` of ``x``.  The only exchange is at set-up: rank 0 evaluates the target function
and assembles the tables once, ``broadcast_layout`` ships them to the other ranks (NCCL over NVLink when the process
group is NCCL, gloo on CPU in the tests), and every rank builds its own device handle from them.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous, balanced partition of ``range(n_rows)``: returns ``(start, stop)`` of this rank's block."""
    base, extra = divmod(int(n_rows), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _device_for_backend():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_layout(layout, src: int = 0):
    """Broadcast a reference-layout dict (NumPy arrays keyed ``offset``, ``F_n``, ``nodes_n`` ..) from ``src``.

    Three collectives regardless of the number of arrays: a header (names, dtypes, shapes), all float64 payloads in
    one buffer, all int64 payloads in one buffer.  Ranks other than ``src`` pass ``None``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return layout
    rank, dev = dist.get_rank(), _device_for_backend()
    header = [None]
    if rank == src:
        keys = sorted(layout)
        header[0] = [(k, "f" if np.asarray(layout[k]).dtype.kind == "f" else "i", tuple(np.asarray(layout[k]).shape)) for k in keys]
    dist.broadcast_object_list(header, src=src)
    spec = header[0]
    out = {}
    for kind, np_dt, t_dt in (("f", np.float64, torch.float64), ("i", np.int64, torch.int64)):
        items = [(k, shp) for k, kd, shp in spec if kd == kind]
        total = int(sum(int(np.prod(shp)) for _, shp in items))
        if total == 0:
            for k, shp in items:
                out[k] = np.zeros(shp, dtype=np_dt)
            continue
        if rank == src:
            flat = np.concatenate([np.ascontiguousarray(layout[k], dtype=np_dt).ravel() for k, _ in items])
            buf = torch.from_numpy(flat).to(dev)
        else:
            buf = torch.empty(total, dtype=t_dt, device=dev)
        dist.broadcast(buf, src=src)
        flat = buf.cpu().numpy()
        pos = 0
        for k, shp in items:
            n = int(np.prod(shp))
            out[k] = flat[pos:pos + n].reshape(shp).copy()
            pos += n
    return out


def max_over_ranks(value: float) -> float:
    """Max of a scalar over ranks (timing is always reported as the slowest rank's device time)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
