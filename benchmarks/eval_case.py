#!/usr/bin/env python
"""One evaluation call of a BASELINE configuration (optionally with another number of outputs), for profiling
(ncu -k regex:fast_wide ...).

    python benchmarks/eval_case.py [cfg4] [points] [d_out]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from smolyax_b200 import _lib, workloads  # noqa: E402
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
base = workloads.CONFIGS[name]
d_out = int(sys.argv[3]) if len(sys.argv) > 3 else base.d_out
wl = workloads.Workload(name, base.rule, base.d_in, d_out, base.n_target, n)
ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=d_out, f=wl.target(), batched_f=True)
x = torch.from_numpy(wl.points(n, seed=3)).cuda()
for _ in range(3):
    y = ip(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
y = ip(x)
b.record()
b.synchronize()
ms = a.elapsed_time(b)
print(f"{name} d_out={d_out} {n} points: {ms:.3f} ms per call = {n * d_out / ms / 1e3:.1f} M evals/s ({_lib.last_kernel()})")
