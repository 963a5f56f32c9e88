#!/usr/bin/env python
"""Block-sparse kernels (K1) against the GEMM-regime kernels (K2) as the number of outputs grows: ms per call and evaluations
per second of both, and what the operator picks by itself (dense=None), on the cfg2 and cfg4 tables.

    python benchmarks/crossover.py [cfg2 cfg4] > profiles/rNN_crossover.txt
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from smolyax_b200 import _lib, workloads  # noqa: E402
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator  # noqa: E402


def timed(fn, reps=5):
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


GRID = {"cfg2": ((1, 2, 3, 4, 6, 8, 10, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64), 200_000), "cfg4": ((1, 2, 4, 6, 8, 10, 12, 16, 20, 24, 32, 40, 48, 64), 100_000)}
for cfg in sys.argv[1:] or ["cfg2", "cfg4"]:
    douts, n = GRID[cfg]
    base = workloads.CONFIGS[cfg]
    for d_out in douts:
        wl = workloads.Workload(cfg, base.rule, base.d_in, d_out, base.n_target, n)
        kw = dict(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=d_out)
        ref = SmolyakBarycentricInterpolator(**kw, f=wl.target(), batched_f=True, dense=False)
        x = torch.from_numpy(wl.points(n, seed=1)).cuda()
        t_sparse = timed(lambda: ref(x))
        k_sparse = _lib.last_kernel()
        alt = SmolyakBarycentricInterpolator(**kw, dense=True)
        alt.set_layout(ref._layout)
        t_dense = timed(lambda: alt(x))
        k_dense = _lib.last_kernel()
        auto = SmolyakBarycentricInterpolator(**kw)
        auto.set_layout(ref._layout)
        t_auto = timed(lambda: auto(x))
        k_auto = _lib.last_kernel()
        err = float((ref(x) - alt(x)).abs().max())
        print(f"{cfg} d_out {d_out:3d}: sparse {t_sparse:8.3f} ms ({k_sparse}) dense {t_dense:8.3f} ms ({k_dense}) "
              f"auto {t_auto:8.3f} ms = {n * d_out / t_auto / 1e3:7.1f} M evals/s ({k_auto})  maxdiff {err:.1e}", flush=True)
        del ref, alt, auto
