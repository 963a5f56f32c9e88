"""Multi-GPU plumbing for the evaluation path: one process per GPU, points partitioned, tables replicated - or, when
``d_out`` is huge, output columns partitioned (``shard_columns`` / ``scatter_columns`` / ``gather_columns``): every rank
sees all points and owns a contiguous block of outputs, i.e. a column slice of the value table.

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


def shard_chunks(n_total: int, chunk: int, rank: int, world: int):
    """Chunks of a long sweep dealt round-robin: yields ``(chunk index, first point, points)`` of this rank's chunks of
    ``range(n_total)`` cut into pieces of ``chunk`` points (the last one may be shorter).  Chunk ``c`` is the same set of points
    at every world size, so a generator keyed by the chunk index produces the same inputs on 1, 2, 4 or 8 GPUs
    (``bench.py``: the 10^8-point sweep of BASELINE's configs[4])."""
    n_chunks = (int(n_total) + int(chunk) - 1) // int(chunk)
    for c in range(int(rank), n_chunks, int(world)):
        first = c * int(chunk)
        yield c, first, min(int(chunk), int(n_total) - first)


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


def shard_columns(d_out: int, rank: int, world: int, align: int = 8):
    """Contiguous partition of the output columns into blocks whose boundaries are multiples of ``align`` (8 = one
    DMMA output block of the GEMM-regime kernel): returns ``(start, stop)`` of this rank's block (may be empty)."""
    blocks = -(-int(d_out) // align)
    lo, hi = shard_rows(blocks, rank, world)
    return min(lo * align, d_out), min(hi * align, d_out)


def column_slice(layout: dict, lo: int, hi: int) -> dict:
    """The tables of outputs ``lo:hi`` only: a layout (compact or reference form) for an interpolator with
    ``d_out = hi - lo``.  The interpolant is linear in ``f``, so the slice evaluates exactly the sliced outputs."""
    out = dict(layout)
    off = np.asarray(layout["offset"], dtype=float).reshape(-1)
    out["offset"] = np.full(hi - lo, off[0]) if off.size == 1 else np.ascontiguousarray(off[lo:hi])
    if layout.get("compact"):
        out["values"] = np.ascontiguousarray(layout["values"][:, lo:hi])
    else:
        for key in layout:
            if key.startswith("F_"):
                out[key] = np.ascontiguousarray(layout[key][:, lo:hi])
    return out


def scatter_columns(layout, d_out: int, src: int = 0) -> dict:
    """Column-sharded set-up for huge ``d_out``: ``src`` holds the full COMPACT layout, every rank receives the index
    arrays (broadcast, small) and only its own ``shard_columns`` slice of the value table (one scatter).
    Ranks other than ``src`` pass ``None``.  Returns the rank's ``column_slice``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return column_slice(layout, 0, d_out)
    rank, world, dev = dist.get_rank(), dist.get_world_size(), _device_for_backend()
    small = None
    if rank == src:
        assert layout.get("compact"), "column sharding ships the compact layout (smx_create_compact)"
        small = {k: v for k, v in layout.items() if k not in ("values", "offset")}
        small["n_values"] = np.asarray(layout["values"].shape[0])
    small = broadcast_layout(small, src=src)
    n_values = int(small.pop("n_values"))
    bounds = [shard_columns(d_out, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in bounds)
    lo, hi = bounds[rank]
    # values and offset travel together: row 0 = offset slice, rows 1.. = value rows, padded to the widest shard
    recv = torch.empty((n_values + 1, max(width, 1)), dtype=torch.float64, device=dev)
    parts = None
    if rank == src:
        off = np.broadcast_to(np.asarray(layout["offset"], dtype=float), (d_out,))
        parts = []
        for a, b in bounds:
            blk = np.zeros((n_values + 1, max(width, 1)))
            blk[0, : b - a] = off[a:b]
            blk[1:, : b - a] = layout["values"][:, a:b]
            parts.append(torch.from_numpy(blk).to(dev))
    dist.scatter(recv, parts, src=src)
    got = recv.cpu().numpy()
    out = dict(small)
    out["compact"] = True
    out["n_active"] = out["n_active"].astype(np.int32)
    out["offset"] = got[0, : hi - lo].copy()
    out["values"] = np.ascontiguousarray(got[1:, : hi - lo])
    return out


def gather_columns(y_local: torch.Tensor, d_out: int) -> torch.Tensor:
    """Replicated ``(N, d_out)`` result from the column shards (optional: the per-call path itself has no collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return y_local
    world = dist.get_world_size()
    bounds = [shard_columns(d_out, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros((y_local.shape[0], width), dtype=y_local.dtype, device=y_local.device)
    pad[:, : y_local.shape[1]] = y_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:, : hi - lo] for p, (lo, hi) in zip(parts, bounds)], dim=1)


def max_over_ranks(value: float) -> float:
    """Max of a scalar over ranks (timing is always reported as the slowest rank's device time)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pin this process (and therefore the page-locked buffers it allocates afterwards: first touch) to the CPUs of the NUMA
    node its GPU hangs off, when the box has more than one node.  Returns what was found, for the bench line.  The end-to-end
    path of several ranks shares the host's PCIe / memory complex; on a box whose GPUs all report the same node there is
    nothing to separate, and the record says so."""
    import os

    out = {"gpu_numa_node": None, "nodes": None, "bound_cpus": None}
    try:
        import torch

        bdf = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        if bdf is None:
            import pynvml as nv

            nv.nvmlInit()
            bdf = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(local_rank)).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:  # NVML prints a 32-bit PCI domain, sysfs a 16-bit one
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        out["gpu_numa_node"], out["nodes"] = node, len(nodes)
        if node >= 0 and len(nodes) > 1 and hasattr(os, "sched_setaffinity"):
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                out["bound_cpus"] = len(cpus)
    except Exception as exc:  # (containers without sysfs access, ...): reported, never fatal
        out["error"] = f"{type(exc).__name__}: {exc}"[:120]
    return out
