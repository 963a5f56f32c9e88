#!/usr/bin/env python
"""A/B timing of K1 (block-sparse value kernel) variants in ONE process: each variant is a set of SMX_* tuning knobs, applied
before the device handle is built.  Needs a tuning build of the library (SMX_TUNING=1 python -m smolyax_b200._build --force);
the product library ignores the environment.

    python benchmarks/k1_ab.py [--config cfg2] [--points N] [--out gpurun_out/k1_ab.jsonl] "A=1 B=2" "A=3" ...

Prints one JSON line per variant: ms per call (CUDA events, median of --reps after 3 warm-ups) and the max relative
difference to the first variant's result on the first 4096 points (ablation variants are wrong on purpose).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from smolyax_b200 import workloads  # noqa: E402
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--d-out", type=int, default=0)
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--out", default="")
    ap.add_argument("variants", nargs="*", default=[""])
    args = ap.parse_args()
    wl = workloads.CONFIGS[args.config]
    if args.d_out:
        wl = workloads.Workload(wl.name, wl.rule, wl.d_in, args.d_out, wl.n_target, wl.n_points)
    ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=wl.d_out, batched_f=True)
    layout = ip._assemble_compact(wl.target(), {})[0]
    gen = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.empty((args.points, wl.d_in), dtype=torch.float64, device="cuda")
    if wl.rule == "leja":
        x.uniform_(-1.0, 1.0, generator=gen)
    else:
        x.normal_(0.0, 2.0 ** -0.5, generator=gen)
    y = torch.empty((args.points, wl.d_out), dtype=torch.float64, device="cuda")
    base = None
    lines = []
    for variant in args.variants:
        knobs = dict(kv.split("=", 1) for kv in variant.split())
        for key in [k for k in os.environ if k.startswith("SMX_") and k != "SMX_TUNING"]:
            del os.environ[key]
        os.environ.update(knobs)
        try:
            ip.set_layout(layout)
            for _ in range(3):
                ip(x, out=y)
            torch.cuda.synchronize()
            times = []
            for _ in range(args.reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ip(x, out=y)
                b.record()
                b.synchronize()
                times.append(a.elapsed_time(b))
            head = y[:4096].cpu().numpy().copy()
            if base is None:
                base = head
            diff = float(np.max(np.abs(head - base)) / np.max(np.abs(base)))
            line = {"variant": variant, "ms": float(np.median(times)), "ms_min": float(min(times)), "diff_vs_first": diff,
                    "config": args.config, "d_out": wl.d_out, "points": args.points, "info": {k: v for k, v in ip.device_info().items()
                                                                          if k in ("n_chunks", "padded_fma", "n_terms")}}
        except Exception as exc:  # a variant the library refuses is reported, the series goes on
            line = {"variant": variant, "error": f"{type(exc).__name__}: {exc}"[:300]}
        print(json.dumps(line), flush=True)
        lines.append(line)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        with open(args.out, "a") as fh:
            for line in lines:
                fh.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
