#!/usr/bin/env python
"""Throughput of every entry point on the BASELINE shapes (bounded sizes), one JSON line per (config, op).

    python benchmarks/sweep.py [--out profiles/rNN_sweep.jsonl]

eval: points*d_out/s with x resident in HBM;  gradient: points*d_out*d_in/s (entries of J);  integral: calls/s.
CUDA events on the launching stream, 3 warm-ups, `--reps` timed repetitions, median reported.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from smolyax_b200 import workloads  # noqa: E402
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator  # noqa: E402

# (config, d_out override, eval points, gradient points)
# (gradient batches are sized to fill the GPU: >= 148 tiles of 32 points, J of at most a few GB)
CASES = [("cfg1", None, 10_000, 10_000), ("cfg2", None, 1_000_000, 37_888), ("cfg3", 64, 100_000, 9_472),
         ("cfg3", None, 100_000, 0), ("cfg4", None, 100_000, 9_472), ("cfg5", None, 1_000_000, 4_736)]
DMMA_PEAK_TFLOPS = 37.12  # profiles/fp64_peaks.json


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        out.append(a.elapsed_time(b))
    return float(np.median(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    lines = []
    for name, d_out, n_eval, n_grad in CASES:
        if args.only and name not in args.only.split(",") and f"{name}:{d_out}" not in args.only.split(","):
            continue
        wl = workloads.CONFIGS[name]
        if d_out is not None:
            wl = workloads.Workload(name, wl.rule, wl.d_in, d_out, wl.n_target, wl.n_points)
        t0 = time.perf_counter()
        ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=wl.d_out, f=wl.target(),
                                            batched_f=True)
        setup_s = time.perf_counter() - t0
        info = ip.device_info()
        gen = torch.Generator(device="cuda").manual_seed(0)
        x = torch.empty((n_eval, wl.d_in), dtype=torch.float64, device="cuda")
        x.uniform_(-1, 1, generator=gen) if wl.rule == "leja" else x.normal_(0, 2 ** -0.5, generator=gen)
        base = {"config": name, "rule": wl.rule, "d_in": wl.d_in, "d_out": wl.d_out, "n_f_evals": ip.n_f_evals,
                "summands": info["n_summands"], "terms": info["n_terms"], "setup_s": round(setup_s, 2)}
        ms = timed(lambda: ip(x), args.reps)
        first = len(lines)
        tf = 2.0 * info["n_terms"] * wl.d_out * n_eval / ms / 1e9
        lines.append({**base, "op": "eval", "kernel": "dense" if info["has_dense_path"] else "sparse", "points": n_eval, "ms": ms,
                      "value": n_eval * wl.d_out / ms * 1e3, "unit": "points*d_out/s",
                      "hbm_gbs_algorithmic": 8.0 * (wl.d_in + wl.d_out) * n_eval / ms / 1e6,
                      "tflops_algorithmic": tf, "frac_of_fp64_dmma_peak": tf / DMMA_PEAK_TFLOPS})
        if info["has_dense_path"] and wl.d_out <= 2048:  # the same tables through the block-sparse kernel, for comparison
            alt = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=wl.d_out, dense=False)
            alt.set_layout(ip._layout)
            xa = x[: max(n_eval // 10, 1000)]
            ms_a = timed(lambda: alt(xa), max(2, args.reps // 2))
            lines.append({**base, "op": "eval", "kernel": "sparse", "points": len(xa), "ms": ms_a,
                          "value": len(xa) * wl.d_out / ms_a * 1e3, "unit": "points*d_out/s"})
            del alt
        if n_grad:
            xg = x[:n_grad]
            ms = timed(lambda: ip.gradient(xg), max(2, args.reps // 2))
            lines.append({**base, "op": "gradient", "kernel": "jobs" if info["grad_jobs"] else "per-summand", "points": n_grad,
                          "ms": ms, "value": n_grad * wl.d_out * wl.d_in / ms * 1e3, "unit": "J entries/s",
                          "points_per_s": n_grad / ms * 1e3, "j_write_gbs": 8.0 * n_grad * wl.d_out * wl.d_in / ms / 1e6})
        ms = timed(lambda: ip.integral(), args.reps)
        lines.append({**base, "op": "integral", "ms": ms, "value": 1e3 / ms, "unit": "calls/s"})
        for ln in lines[first:]:
            print(json.dumps(ln), flush=True)
        del ip, x
        torch.cuda.empty_cache()
    if args.out:
        Path(args.out).write_text("".join(json.dumps(ln) + "\n" for ln in lines))


if __name__ == "__main__":
    main()
