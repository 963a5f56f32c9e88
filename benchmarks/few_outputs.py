#!/usr/bin/env python
"""2..8 outputs on the cfg2 tables: ms per 1e6 points of the block-sparse path (which kernel serves them is decided
at create time; SMX_FAST_MULTI=0 forces one output per pass of the lean kernel)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from smolyax_b200 import workloads  # noqa: E402
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator  # noqa: E402

wl = workloads.CONFIGS["cfg2"]
x = torch.rand((500_000, wl.d_in), dtype=torch.float64, device="cuda") * 2 - 1
for d_out in (1, 2, 3, 4, 6, 8):
    w = workloads.Workload("cfg2", wl.rule, wl.d_in, d_out, wl.n_target, wl.n_points)
    ip = SmolyakBarycentricInterpolator(node_gen=w.generator(), k=w.k(), t=w.threshold(), d_out=d_out, f=w.target(), batched_f=True)
    for _ in range(3):
        ip(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ip(x)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / 5 * 2  # per 1e6 points
    print(json.dumps({"d_out": d_out, "ms_per_1e6_points": ms, "ms_per_output": ms / d_out}), flush=True)
