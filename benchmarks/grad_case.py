#!/usr/bin/env python
"""One gradient call of a BASELINE configuration, for profiling (ncu -k regex:grad_kernel ...).

    python benchmarks/grad_case.py [cfg4] [points]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from smolyax_b200 import workloads  # noqa: E402
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
wl = workloads.CONFIGS[name]
ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=wl.d_out, f=wl.target(), batched_f=True)
x = torch.from_numpy(wl.points(n, seed=3)).cuda()
for _ in range(3):
    J = ip.gradient(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
J = ip.gradient(x)
b.record()
b.synchronize()
print(name, n, "points:", a.elapsed_time(b), "ms per gradient call", ip.device_info()["grad_jobs"], "jobs", ip.device_info()["grad_items"], "items")
