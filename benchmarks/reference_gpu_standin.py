"""Stand-in for "the reference's JAX-on-GPU path" (BASELINE north_star: an additional reported number).

JAX is not installable in this image, so the reference cannot run on the GPU itself.  This module restates the
reference's *algorithm and batching* with stock torch fp64 ops on the same B200 - NOT this package's kernels, NOT JAX/XLA:

  * per group n and per memory-limited batch of summands (reference interpolation.py:283-290, same formula, same
    default of 4 GB) one vectorised call over the batch, the twin of jit(vmap(evaluate_tensor_product_interpolant))
    (barycentric.py:69-123): gather the active columns of x (:116), un-normalised basis w / (x - xi) with the degree
    mask and the one-hot pattern at node hits (:60-66), normalisation by the row sum (:117), one einsum against the
    zero-padded value tensors (:119-123), times zeta, summed over the batch (interpolation.py:302);
  * the (S, N, m) bases and the einsum intermediates are materialised in HBM, as XLA's un-fused dot_general chain does.

It is a benchmark baseline only: nothing under smolyax_b200/ imports it.  Usage (on a GPU box):

    python benchmarks/reference_gpu_standin.py --config cfg2 --points 100000
"""
from __future__ import annotations

import argparse
import json
import string
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def upload(layout: dict, device="cuda") -> dict:
    """The six per-group arrays of the reference layout as device tensors (interpolation.py:230-235)."""
    out = {"offset": torch.as_tensor(np.asarray(layout["offset"], dtype=np.float64), device=device), "groups": []}
    for n in sorted(int(key.split("_")[1]) for key in layout if key.startswith("zetas_")):
        g = {"n": n}
        for key, dt in (("F", torch.float64), ("nodes", torch.float64), ("weights", torch.float64), ("dims", torch.int64),
                        ("degs", torch.int64), ("zetas", torch.int64)):
            g[key] = torch.as_tensor(np.ascontiguousarray(layout[f"{key}_{n}"]), device=device).to(dt)
        out["groups"].append(g)
    return out


def _batch(x, F, xi, w, dims, degs, zetas):
    """sum_s zeta_s I_s(x) for one batch: F (S, d_out, m_1..m_n), xi / w (S, n, m_max), dims / degs (S, n)."""
    S, n = dims.shape
    bs = []
    for j in range(n):
        m = F.shape[2 + j]
        xj = x[:, dims[:, j]].T[:, :, None]                      # (S, N, 1)   x[:, [s_j]]
        diffs = xj - xi[:, j, None, :m]                           # (S, N, m)
        cols = torch.arange(m, device=x.device)[None, None, :] <= degs[:, j, None, None]
        hit = (diffs == 0) & cols
        row_hit = hit.any(dim=2, keepdim=True)
        b = torch.where(row_hit, (diffs == 0).to(x.dtype), w[:, j, None, :m] / diffs)
        b = torch.where(cols, b, torch.zeros((), dtype=x.dtype, device=x.device))
        bs.append(b / b.sum(dim=2, keepdim=True))
    ax = string.ascii_lowercase
    f_axes = ax[: n + 1]                                          # o, m_1..m_n
    spec = "S" + f_axes + "," + ",".join("SN" + a for a in f_axes[1:]) + "->SN" + f_axes[0]
    res = torch.einsum(spec, F, *bs)                              # (S, N, d_out)
    return (zetas.to(x.dtype)[:, None, None] * res).sum(dim=0)


def evaluate(tables: dict, x: torch.Tensor, memory_limit: float = 4.0) -> torch.Tensor:
    n_points, d_out = x.shape[0], tables["groups"][0]["F"].shape[1] if tables["groups"] else tables["offset"].numel()
    y = tables["offset"].expand(n_points, d_out).clone()
    for g in tables["groups"]:
        F = g["F"]
        per_summand = n_points * d_out * float(np.prod(F.shape[3:])) * 8 / 1024 ** 3
        step = max(1, int(np.floor(memory_limit / per_summand)))
        for s in range(0, F.shape[0], step):
            e = min(s + step, F.shape[0])
            y += _batch(x, F[s:e], g["nodes"][s:e], g["weights"][s:e], g["dims"][s:e], g["degs"][s:e], g["zetas"][s:e])
    return y


def timed(layout: dict, x: torch.Tensor, repeats: int = 2, memory_limit: float = 4.0):
    """(points/s, seconds of the best repeat, y) - CUDA events, one warm-up call."""
    tables = upload(layout, x.device)
    y = evaluate(tables, x, memory_limit)
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(repeats):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        y = evaluate(tables, x, memory_limit)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return x.shape[0] / best, best, y


def main():
    from smolyax_b200 import workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--points", type=int, default=100_000)
    ap.add_argument("--memory-limit", type=float, default=4.0)
    args = ap.parse_args()
    wl = workloads.CONFIGS[args.config]
    ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=wl.d_out, batched_f=True,
                                        layout="reference", f=wl.target())
    x = torch.from_numpy(wl.points(args.points, seed=7)).cuda()
    t0 = time.perf_counter()
    pps, secs, y = timed(ip.reference_layout(), x, memory_limit=args.memory_limit)
    ours = ip(x)
    err = float((y - ours).abs().max() / ours.abs().max())
    print(json.dumps({"config": args.config, "points": args.points, "standin_points_per_s": pps, "standin_s": secs,
                      "evals_per_s": pps * wl.d_out, "max_diff_vs_b200_kernels_rel_to_max": err,
                      "wall_s": time.perf_counter() - t0,
                      "what": "torch fp64 eager restatement of the reference's vmap/einsum batches on this GPU (not JAX)"}))


if __name__ == "__main__":
    main()
