#!/usr/bin/env python
"""Shared-memory wavefronts of the A-fragment factor loads, from the planner's own model (no GPU needed).

    python benchmarks/bank_stats.py cfg2 [cfg4 ...]
"""
import ctypes
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import test_plan as tp  # noqa: E402
from helpers import load  # noqa: E402

tp._host.smxh_plan_bank_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
for case in sys.argv[1:] or ["cfg2"]:
    g = load(case)
    layout = tp._layout_of(g, case)
    t0 = time.perf_counter()
    plan = tp.Plan(layout, g["x"].shape[1], int(g["d_out"]), options=tp.SPARSE)
    dt = time.perf_counter() - t0
    st = np.zeros(6, dtype=np.int64)
    tp._host.smxh_plan_bank_stats(plan.h, st.ctypes.data)
    loads, wf, wf_real, ks, halves, halves_min = st.tolist()
    print(f"{case}: plan {dt:.2f} s, items {plan.stats['n_chunks']}, k-steps {ks}, factor loads {loads}, wavefront groups {wf} "
          f"(x{wf / loads:.3f}), without the ones row {wf_real} (x{wf_real / loads:.3f}); (k-step, half block) pairs {halves}, "
          f"fewest possible with these items {halves_min}; terms {plan.stats['n_terms']}, padded FMAs {plan.stats['padded_fma']}")
