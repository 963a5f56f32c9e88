#!/usr/bin/env python
"""Dense (GEMM-regime) kernel at a few output widths on the cfg2 tables (mostly cold leading entries) and the cfg3 tables
(all hot): fraction of the FP64 DMMA rate.  Kernel shape knobs come from the environment (SMX_DENSE_SPLITK, SMX_DENSE_NB ..)."""
import sys
import numpy as np
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from smolyax_b200 import workloads
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        out.append(a.elapsed_time(b))
    return float(np.median(out))


for cfg, n in (("cfg2", 200_000), ("cfg3", 100_000)):
    base = workloads.CONFIGS[cfg]
    for d_out in (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "16,32,64,100,128").split(",")):
        wl = workloads.Workload(cfg, base.rule, base.d_in, d_out, base.n_target, n)
        ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=d_out, f=wl.target(), batched_f=True, dense=True)
        x = torch.from_numpy(wl.points(n, seed=1)).cuda()
        ms = timed(lambda: ip(x))
        terms = ip.device_info()["n_terms"]
        print(f"{cfg} d_out {d_out:4d}  {ms:8.3f} ms  {n * d_out / ms / 1e6:8.1f} Mevals/s  frac {2.0 * terms * d_out * n / ms / 1e9 / 37.12:.3f}", flush=True)
        del ip
