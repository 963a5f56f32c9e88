#!/usr/bin/env python
"""The reference's heat-map grid (benchmarking/benchmark.py:33-36, 104-136, 187-222) on the B200 path.

The reference times, for every cell of
    d_in in {10, 40, 160} x d_out in {10, 40, 160} x n_batch in {50, 500, 5000} x |Lambda| in {1e3 .. 1e4},
``N_ITER = 20`` calls ``np.asarray(interp(X))`` from host NumPy to host NumPy (wall clock), with the target family of
benchmarking/testfunction.py and the anisotropy ``k_j = log((j + 2)^r / theta)`` (benchmark.py:180), and reports the
relative RMSE against ``f`` (benchmark.py:66).  Its opponent there is Tasmanian, which this image does not have.  This
script reports the B200 path alone; ``python bench.py --heatmap`` runs the same grid WITH a baseline column (the CPU
restatement of the reference's algorithm on the same X, which only bench.py may execute) and prints
log10(t_cpu / t_b200) panels like benchmark.py:190-222.

    python benchmarks/heatmap.py [--out profiles/rNN_heatmap.jsonl] [--quick] [--cache benchmarks/results]

One JSON line per cell; a text table of per-call milliseconds per (d_in, d_out) panel at the end.
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
from smolyax_b200 import indices, nodes, workloads  # noqa: E402
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator  # noqa: E402

TARGET_N_LIST = [1_000, 2_000, 4_000, 6_000, 8_000, 10_000]  # benchmark.py:33
N_MC_LIST = [50, 500, 5000]                                   # benchmark.py:34
D_IN_LIST = [10, 40, 160]                                     # benchmark.py:35
D_OUT_LIST = [10, 40, 160]                                    # benchmark.py:36
N_ITER = 20                                                   # benchmark.py:38


def rmse(y_true, y_pred):  # benchmark.py:66
    return float(np.sqrt(np.mean((y_true - y_pred) ** 2) / np.mean(y_true ** 2)))


def run_grid(quick=False, seed=0, cpu_timer=None, cache_dir=None, use_saved=True, log=print):
    """All cells of the grid.  ``cpu_timer(layout, X) -> (seconds per call, note)`` adds the baseline column (bench.py passes the
    CPU oracle; this module never imports it).  ``cache_dir``: one .npz of timings per cell, reused when present, as the
    reference does (benchmark.py:86-88, 119-121)."""
    assert torch.cuda.is_available(), "needs a GPU: there is no CPU path"
    d_ins, d_outs, n_mcs, n_targets = D_IN_LIST, D_OUT_LIST, N_MC_LIST, TARGET_N_LIST
    if quick:
        d_ins, d_outs, n_mcs, n_targets = [10, 160], [10, 160], [50, 5000], [1_000, 10_000]
    if cache_dir is not None:
        Path(cache_dir).mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(seed)
    lines = []
    for d_in in d_ins:
        # anisotropy of the reference's benchmark (benchmark.py:180): k_j = log((j + 2)^r / theta)
        k = np.array([np.log(((j + 2) ** workloads.BASE_R) / workloads.BASE_THETA) for j in range(d_in)])
        node_gen = nodes.Leja(dim=d_in)
        for d_out in d_outs:
            f = workloads.TargetFamily(d_in, d_out)
            X_full = rng.uniform(-1.0, 1.0, size=(max(N_MC_LIST), d_in))  # benchmark.py:187
            Y_full = f(X_full)
            for n_target in n_targets:
                cached = {}
                for n_mc in n_mcs:
                    fname = None if cache_dir is None else Path(cache_dir) / (
                        f"b200_din{d_in}_dout{d_out}_N{n_target}_nmc{n_mc}_niter{N_ITER}_seed{seed}{'_cpu' if cpu_timer else ''}.npz")
                    if fname is not None and fname.is_file() and use_saved:
                        cached[n_mc] = json.loads(str(np.load(fname)["line"][()]))
                if len(cached) == len(n_mcs):
                    for n_mc in n_mcs:
                        lines.append(cached[n_mc])
                        log(json.dumps(cached[n_mc]))
                    continue
                t_thr = indices.find_approximate_threshold(k, n_target, node_gen.is_nested)
                t0 = time.perf_counter()
                ip = SmolyakBarycentricInterpolator(node_gen=node_gen, k=k, t=t_thr, d_out=d_out, f=f, n_inputs=max(n_mcs),
                                                    batched_f=True, layout="reference" if cpu_timer else "auto")
                setup_s = time.perf_counter() - t0
                info = ip.device_info()
                for n_mc in n_mcs:
                    X = np.ascontiguousarray(X_full[:n_mc])
                    Y = np.asarray(ip(X))  # warm-up (the reference's first call compiles)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(N_ITER):
                        Y = np.asarray(ip(X))  # benchmark.py:131
                    runtime = time.perf_counter() - t0
                    line = {"d_in": d_in, "d_out": d_out, "n_target": n_target, "n_f_evals": ip.n_f_evals, "n_batch": n_mc,
                            "n_iter": N_ITER, "runtime_s": runtime, "ms_per_call": 1e3 * runtime / N_ITER,
                            "evals_per_s": n_mc * d_out * N_ITER / runtime, "rmse": rmse(Y_full[:n_mc], Y),
                            "setup_s": setup_s, "path": "dense" if info["has_dense_path"] else "block-sparse"}
                    if cpu_timer is not None:
                        cpu_s, note = cpu_timer(ip.reference_layout(), X)
                        line.update(cpu_ms_per_call=1e3 * cpu_s, cpu_note=note, log10_cpu_over_b200=float(np.log10(cpu_s * N_ITER / runtime)))
                    lines.append(line)
                    log(json.dumps(line))
                    if cache_dir is not None:
                        np.savez_compressed(Path(cache_dir) / (
                            f"b200_din{d_in}_dout{d_out}_N{n_target}_nmc{n_mc}_niter{N_ITER}_seed{seed}{'_cpu' if cpu_timer else ''}.npz"),
                            line=np.array(json.dumps(line)))
                del ip
    return lines, (d_ins, d_outs, n_mcs, n_targets)


def print_panels(lines, grid, key="ms_per_call", title="ms per call", fmt="{:8.3f}", file=sys.stderr):
    """Text panels like the reference's heat maps (benchmark.py:190-222): rows n_batch, columns |Lambda|."""
    d_ins, d_outs, n_mcs, n_targets = grid
    for d_in in d_ins:
        for d_out in d_outs:
            print(f"\n# d_in={d_in} d_out={d_out}: {title}; columns n_target = {n_targets}", file=file)
            for n_mc in n_mcs:
                row = [next(l[key] for l in lines if (l["d_in"], l["d_out"], l["n_target"], l["n_batch"]) == (d_in, d_out, nt, n_mc))
                       for nt in n_targets]
                print(f"  n_batch={n_mc:5d}: " + " ".join(fmt.format(v) for v in row), file=file)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--quick", action="store_true", help="corners of the grid only")
    ap.add_argument("--cache", default="", help="directory for per-cell .npz timings (reused when present)")
    args = ap.parse_args()
    lines, grid = run_grid(quick=args.quick, seed=args.seed, cache_dir=args.cache or None)
    if args.out:
        Path(args.out).write_text("".join(json.dumps(l) + "\n" for l in lines))
    print_panels(lines, grid)


if __name__ == "__main__":
    main()
