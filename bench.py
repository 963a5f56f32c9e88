#!/usr/bin/env python
"""Benchmark of the hot path: batched evaluation of the barycentric Smolyak interpolant.

    python bench.py --gpus N --steps K --warmup W            # this framework (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm (CPU oracle port) on host cores

Workload of the headline line (BASELINE.json configs[1], the configuration the metric is quoted on): Leja nodes, d_in = 1000,
d_out = 1, n = 10^4 nodes, 10^6 evaluation points per GPU (weak scaling), synthetic inputs U(-1,1)^d, fp64.
A step = one `__call__` over the whole batch = ONE fused kernel launch.  `value` is measured with the batch resident in HBM
(8 GB of x per step, far larger than the 126 MB L2, so no flush is needed); `e2e` is the same batch evaluated through the
public API from pinned HOST memory, copies in and out inside the timed region.

The same JSON line carries, under `others`, one short record per remaining BASELINE configuration and entry point (cfg1,
cfg3 at full size, cfg4, cfg5: `__call__`, `gradient`, `integral`) with its time, its roofline and its parity in both norms
and against the 80-bit referee, and `cfg5_sweep`: BASELINE configs[4], 10^8 points at d_out = 100 sharded over the ranks
(strong scaling, x generated on the device chunk by chunk).  Prints one JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "interpolant evals/sec (points*d_out/s) at d_in=1e3,n=1e4"
UNIT = "points*d_out/s"


_RESULT_STREAM = None


def claim_stdout():
    """stdout carries the ONE JSON line and nothing else: whatever a library prints to file descriptor 1 during the run
    (NCCL's version banner when NCCL_DEBUG is set, as on the GPU boxes) is sent to stderr instead."""
    global _RESULT_STREAM
    sys.stdout.flush()
    _RESULT_STREAM = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    print(json.dumps(line), file=_RESULT_STREAM or sys.stdout, flush=True)


def measured_peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return json.loads(path.read_text()), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def fp64_peak():
    """FP64 tensor (DMMA) rate measured on this GPU type with profiles/fp64_peaks.cu: MEASURED_PEAKS.json only carries the
    bf16 tensor rate, which no fp64 path can use."""
    try:
        return float(json.loads((ROOT / "profiles" / "fp64_peaks.json").read_text())["fp64_dmma_tflops"])
    except Exception:
        return 37.12


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs (NVML is initialised up front so
    that the first sample falls inside the region)."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.nv, self.handle = [], set(), None, None, None
        self._stop_evt = threading.Event()
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv, self.handle = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self._sample()  # warm the NVML path; discarded below
            self.samples.clear()
        except Exception as exc:  # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def _sample(self):
        nv, h = self.nv, self.handle
        self.samples.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self._sample()
            except Exception as exc:
                self.reasons.add(f"nvml_error:{type(exc).__name__}")
                return

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


ORACLE_COLUMNS = 64  # outputs the CPU oracle evaluates when the padded reference layout of all d_out would not fit


def oracle_layout(wl):
    """Reference layout (the oracle's input) of the workload, assembled on the host (no GPU, no CUDA library); for huge d_out
    restricted to ORACLE_COLUMNS evenly spaced output columns (the padded F_n of all 10^4 outputs of cfg3 would be 29 GB).
    Returns (layout, columns or None)."""
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    cols = None
    fam = wl.target()
    f = fam
    if wl.d_out > 256:
        cols = np.unique(np.linspace(0, wl.d_out - 1, ORACLE_COLUMNS).astype(np.int64))
        f = lambda x: fam(x)[..., cols]  # noqa: E731
    ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), layout="reference", batched_f=True,
                                        d_out=wl.d_out if cols is None else len(cols))
    return ip._assemble(f, {})[0], cols


def cpu_reference(wl, layout, seconds: float, repeats: int, cols=None):
    """Times the oracle (plain-C restatement of the reference algorithm, OpenMP over points) on the host cores on a
    bounded sample of the workload.  Returns (points*d_out/s, cores, sample description, per-repeat seconds)."""
    from oracle import oracle

    cores = oracle.use_all_cores()
    d_eff = wl.d_out if cols is None else len(cols)
    x = wl.points(max(cores * 4, 32), seed=123)
    t0 = time.perf_counter()
    oracle.evaluate(layout, x)
    per_point = (time.perf_counter() - t0) / len(x)
    n = int(min(max(seconds / max(per_point, 1e-9), cores), 200_000))
    x = wl.points(n, seed=124)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        oracle.evaluate(layout, x)
        times.append(time.perf_counter() - t0)
    best = min(times)
    what = "the same workload" if cols is None else f"the same workload restricted to {d_eff} of its {wl.d_out} outputs"
    return n * d_eff / best, cores, f"{n} points of {what}, best of {repeats} ({best:.2f} s each)", times


def run_reference(args):
    """Reference arm: the reference's algorithm on the host cores.  Every step is a bounded sample of the workload (about
    --cpu-seconds of CPU work); `steps` in the line is the number of samples really timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from smolyax_b200 import workloads

    wl = workloads.CONFIGS[args.config]
    layout, cols = oracle_layout(wl)
    steps = max(1, min(args.steps, 3))  # each sample is ~args.cpu_seconds of all-core work: three are plenty
    warmup = min(args.warmup, 1)
    for _ in range(warmup):
        cpu_reference(wl, layout, 0.5, 1, cols)
    value, cores, sample, times = cpu_reference(wl, layout, args.cpu_seconds, steps, cols)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 * min(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: {wl.rule} d_in={wl.d_in} d_out={wl.d_out} n={wl.n_target}", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference algorithm (padded per-summand second-barycentric-form contraction) restated in C + OpenMP "
                "(oracle/smx_oracle.c); the reference's JAX runtime is not installable offline.  A step is a bounded sample of "
                "the workload, best of `steps` samples; the tables are assembled on the host, the CUDA library is not loaded",
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------------------
# parity of a device result against the CPU oracle, in both norms, and of both against the 80-bit referee
# ------------------------------------------------------------------------------------------------------------------------
def value_parity(layout, xs, got):
    """got (n, d_eff) against the oracle on the same points: pointwise relative error; error relative to the summand magnitude
    sum_nu |zeta_nu I_nu f| (the norm the 1e-12 bound is stated in: the signed Smolyak sum carries sum|zeta| ~ 1e4 times more
    rounding noise than its value suggests); and the distance of this implementation and of the fp64 reference algorithm
    from the same algorithm in 80-bit arithmetic."""
    from oracle import oracle

    ref = oracle.evaluate(layout, xs)
    mag_layout = {k: (np.abs(v) if k.startswith("zetas_") or k == "offset" else v) for k, v in layout.items()}
    mag = oracle.evaluate(mag_layout, xs)
    ld = oracle.evaluate_referee(layout, xs)
    scale = float(np.max(np.abs(ref)))
    return {
        "points": int(len(xs)),
        "max_rel_vs_oracle": float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))),
        "scaled_by_summand_magnitude": float(np.max(np.abs(got - ref) / np.maximum(np.abs(mag), 1e-300))),
        "this_vs_80bit_referee": float(np.max(np.abs((got - ld).astype(np.float64))) / scale),
        "fp64_reference_algorithm_vs_80bit_referee": float(np.max(np.abs((ref - ld).astype(np.float64))) / scale),
    }


def gradient_parity(layout, xs, got):
    from oracle import oracle

    ref = oracle.gradient(layout, xs)
    ld = oracle.gradient_referee(layout, xs)
    ok = ~np.isnan(ref)
    scale = max(float(np.max(np.abs(ref[ok]))), 1e-300) if ok.any() else 1.0
    return {
        "points": int(len(xs)),
        "nan_pattern_equal": bool(np.array_equal(np.isnan(got), np.isnan(ref))),
        "max_abs_vs_oracle_rel_to_max": float(np.max(np.abs((got - ref)[ok]), initial=0.0) / scale),
        "this_vs_80bit_referee": float(np.max(np.abs((got - ld)[ok].astype(np.float64)), initial=0.0) / scale),
        "fp64_reference_algorithm_vs_80bit_referee": float(np.max(np.abs((ref - ld)[ok].astype(np.float64)), initial=0.0) / scale),
    }


def eval_roofline(info, d_in, d_loc, n_points, ms, peaks, peak_kind, launches_per_step=None, traffic=None, traffic_source=None, kernel_name=None):
    """t_roof = max(bytes / BW, flops / FP64 rate) with bytes_eval = 8 (d_in + d_out) per point and the folded
    form's 2 * n_terms * d_out flops per point (padding is not counted)."""
    pk = fp64_peak()
    alg_bytes = 8.0 * (d_in + d_loc) * n_points
    alg_flops = 2.0 * info["n_terms"] * d_loc * n_points
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    tf = alg_flops / (ms * 1e-3) / 1e12
    dense = bool(info["has_dense_path"])
    fma_per_eval = info["dense_terms"] if dense else info["padded_fma"]
    tensor_bound = dense or alg_flops / (pk * 1e12) > alg_bytes / (peaks["hbm_gbs"] * 1e9)
    # (smx_last_kernel(): the instantiation that really ran; the fallback names the family the dispatch rules would pick)
    kernel = kernel_name or ("dense_eval_kernel" if dense else ("fast_pipe_kernel" if d_loc == 1 else "fast_multi_kernel" if d_loc in (3, 5, 6) else "fast_lean_kernel"))
    if tensor_bound:
        r = {"bound": "tensor", "achieved": tf, "peak": pk, "unit": "TFLOP/s", "frac": tf / pk, "traffic": traffic,
             "peak_source": "FP64 DMMA (mma.sync.m8n8k4.f64) rate measured with profiles/fp64_peaks.cu (MEASURED_PEAKS.json has no fp64 figure)",
             "kernel": kernel, "algorithmic_flops_per_launch": alg_flops, "hbm_gbs_algorithmic": gbs, "hbm_peak_gbs": peaks["hbm_gbs"]}
    else:
        r = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": traffic,
             "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)", "kernel": kernel, "algorithmic_bytes_per_launch": alg_bytes,
             "fp64_dmma_peak_tflops_measured": pk,
             "note": f"x is streamed once (8*(d_in+d_out) B per point): {alg_bytes / (peaks['hbm_gbs'] * 1e9) * 1e3:.2f} ms per launch at the measured HBM "
                     f"peak; the FP64 tensor work it executes (2*padded_fma flop per point) needs "
                     f"{2.0 * fma_per_eval * d_loc * n_points / (pk * 1e12) * 1e3:.2f} ms at the measured DMMA peak"}
    r["fp64_tflops_executed"] = 2.0 * fma_per_eval * d_loc * n_points / (ms * 1e-3) / 1e12
    if traffic_source:
        r["traffic_source"] = traffic_source
    if launches_per_step is not None:
        r["launches_per_step"] = launches_per_step
    if ms < 0.2:
        r["note"] = "a step of this size is launch/latency-bound (tens of microseconds per call), not pipe-bound"
    return r


def gradient_roofline(wl, n_points, ms, peaks):
    """bytes_grad = 8 (d_in + d_out d_in) per point (J is written once); folded algorithmic flops
    2 * d_out * sum over terms of their active dimensions per point."""
    from smolyax_b200 import indices

    pk = fp64_peak()
    incidences = sum(len(nu) for nu in indices.indexset(wl.k(), wl.threshold()))
    alg_bytes = 8.0 * (wl.d_in + wl.d_out * wl.d_in) * n_points
    alg_flops = 2.0 * wl.d_out * incidences * n_points
    t_bytes, t_flops = alg_bytes / (peaks["hbm_gbs"] * 1e9), alg_flops / (pk * 1e12)
    bound = "hbm" if t_bytes >= t_flops else "tensor"
    achieved = alg_bytes / (ms * 1e-3) / 1e9 if bound == "hbm" else alg_flops / (ms * 1e-3) / 1e12
    peak = peaks["hbm_gbs"] if bound == "hbm" else pk
    return {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": achieved / peak,
            "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_flops_per_launch": alg_flops,
            "j_write_gbs": 8.0 * wl.d_out * wl.d_in * n_points / (ms * 1e-3) / 1e9}


def timed_ms(fn, reps, warm=2):
    import torch

    for _ in range(warm):
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


# (config, eval points, gradient points): gradient batches fill the GPU (>= 148 tiles of 32 points) with J of <= 4 GB
OTHER_CASES = [("cfg1", 10_000, 10_000), ("cfg3", 100_000, 0), ("cfg4", 100_000, 9_472), ("cfg5", 1_000_000, 4_736)]


def other_records(peaks, peak_kind, device):
    """One record per remaining BASELINE configuration and entry point (rank 0, one GPU)."""
    import torch

    from smolyax_b200 import _lib, workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    records = []
    for name, n_eval, n_grad in OTHER_CASES:
        try:
            wl = workloads.CONFIGS[name]
            t0 = time.perf_counter()
            compact = wl.d_out > 256
            ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=wl.d_out, device=device,
                                                batched_f=True, layout="compact" if compact else "reference")
            ip.set_f(f=wl.target())
            setup_s = time.perf_counter() - t0
            info = ip.device_info()
            check_layout, cols = (ip.reference_layout(), None) if not compact else oracle_layout(wl)
            gen = torch.Generator(device="cuda").manual_seed(4321)
            x = torch.empty((n_eval, wl.d_in), dtype=torch.float64, device="cuda")
            x.uniform_(-1, 1, generator=gen) if wl.rule == "leja" else x.normal_(0, 2 ** -0.5, generator=gen)
            y = torch.empty((n_eval, wl.d_out), dtype=torch.float64, device="cuda")
            base = {"config": name, "workload": f"{wl.rule} d_in={wl.d_in} d_out={wl.d_out} n={wl.n_target} ({info['n_terms']} terms)",
                    "setup_s": round(setup_s, 2)}
            l0 = _lib.lib.smx_launch_count()
            ms = timed_ms(lambda: ip(x, out=y), 3 if name == "cfg3" else 5, warm=1 if name == "cfg3" else 2)
            launches = int(_lib.lib.smx_launch_count() - l0) // ((3 if name == "cfg3" else 5) + (1 if name == "cfg3" else 2))
            kernel_eval = _lib.last_kernel()
            got = y[:64].cpu().numpy()
            if cols is not None:
                got = got[:, cols]
            records.append({**base, "op": "eval", "points": n_eval, "ms": ms, "value": n_eval * wl.d_out / (ms * 1e-3), "unit": UNIT,
                            "kernel": kernel_eval, "launches_per_call": launches,
                            "roofline": eval_roofline(info, wl.d_in, wl.d_out, n_eval, ms, peaks, peak_kind, launches, kernel_name=kernel_eval),
                            "parity": value_parity(check_layout, x[:64].cpu().numpy(), got)})
            if n_grad:
                xg = x[:n_grad]
                ms = timed_ms(lambda: ip.gradient(xg), 3, warm=1)
                kernel_grad = _lib.last_kernel()
                J = ip.gradient(xg[:8]).cpu().numpy()
                records.append({**base, "op": "gradient", "points": n_grad, "ms": ms, "value": n_grad * wl.d_out * wl.d_in / (ms * 1e-3),
                                "unit": "J entries/s", "points_per_s": n_grad / (ms * 1e-3),
                                "kernel": kernel_grad,
                                "roofline": gradient_roofline(wl, n_grad, ms, peaks),
                                "parity": gradient_parity(check_layout, xg[:8].cpu().numpy(), J)})
            ms = timed_ms(lambda: ip.integral(), 5)
            rec = {**base, "op": "integral", "ms": ms,
                   "roofline": {"bound": "hbm", "note": "one pass over the value tensors (8*d_out*W bytes; microseconds): launch/latency-bound"}}
            if cols is None:
                from oracle import oracle

                q_ref = oracle.integral(check_layout)
                rec["parity"] = {"max_abs_vs_oracle_rel_to_max": float(np.max(np.abs(ip.integral() - q_ref)) / max(1e-300, float(np.max(np.abs(q_ref)))))}
            records.append(rec)
            del ip, x, y
            torch.cuda.empty_cache()
        except Exception as exc:  # never fatal for the headline line
            records.append({"config": name, "error": f"{type(exc).__name__}: {exc}"[:300]})
    return records


def cfg5_sweep(args, world, rank, local, peaks):
    """BASELINE configs[4]: 10^8 evaluation points at d_in = 10^3, n = 10^4, d_out = 100, sharded over the ranks (strong scaling;
    no data-path collective).  x (800 GB in all) is generated on the device chunk by chunk by torch's counter-based (Philox)
    generator keyed (seed, chunk); only the evaluations are timed (CUDA events around each call, summed)."""
    import torch
    import torch.distributed as dist

    from smolyax_b200 import dist as sdist, workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    wl = workloads.CONFIGS["cfg5"]
    total = args.sweep_points
    chunk = min(500_000, total)  # 200 chunks at the default total: the same number of chunks per rank at 1, 2, 4 and 8 GPUs
    ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=wl.d_out, device=local, batched_f=True)
    layout = ip._assemble(wl.target(), {})[0] if rank == 0 else None
    layout = sdist.broadcast_layout(layout, src=0)
    ip.set_layout(layout)
    info = ip.device_info()
    x = torch.empty((chunk, wl.d_in), dtype=torch.float64, device="cuda")
    y = torch.empty((chunk, wl.d_out), dtype=torch.float64, device="cuda")
    gen = torch.Generator(device="cuda")
    for _ in range(2):
        ip(x.uniform_(-1, 1), out=y)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, points, checksum = 0.0, 0, 0.0
    t_wall = time.perf_counter()
    for c, _, n in sdist.shard_chunks(total, chunk, rank, world):
        gen.manual_seed(977 * 1_000_003 + c)
        x.uniform_(-1, 1, generator=gen)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ip(x[:n], out=y[:n])
        b.record()
        b.synchronize()
        ms += a.elapsed_time(b)
        points += n
        if c == 0:
            checksum = float(y[:n].sum())  # (chunk 0 is rank 0's at every world size: same number at 1, 2, 4 and 8 GPUs)
    wall = time.perf_counter() - t_wall
    ms_max = sdist.max_over_ranks(ms)
    wall_max = sdist.max_over_ranks(wall)
    pk = fp64_peak()
    flops = 2.0 * info["n_terms"] * wl.d_out * points  # this rank's share
    return {
        "workload": f"cfg5: leja d_in={wl.d_in} d_out={wl.d_out} n={wl.n_target}, {total} points in chunks of {chunk}, sharded over {world} GPU(s)",
        "scaling": "strong", "n_gpus": world, "points_total": total, "points_per_gpu": points, "kernel_ms_max_over_ranks": ms_max,
        "value": total * wl.d_out / (ms_max * 1e-3), "unit": UNIT,
        "wall_s_incl_generation": wall_max, "value_incl_generation": total * wl.d_out / wall_max,
        "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": pk, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / pk,
                     "kernel": "dense_eval_kernel", "note": "rank 0's share; algorithmic 2*n_terms*d_out flop per point on the FP64 tensor instruction"},
        "data": "synthetic: U(-1,1), generated on the device per chunk (Philox, keyed by chunk), generation outside the timed region",
        "checksum_chunk0": checksum,
    }


def run_heatmap(args):
    """the reference's benchmark grid (benchmarking/benchmark.py:33-36, 104-136) with a baseline column - per
    cell the B200 path as the reference times its own (20 calls host NumPy -> host NumPy) and the CPU restatement of the
    reference's algorithm (oracle, all host cores) on the same X; log10(t_cpu / t_b200) panels like benchmark.py:190-222.
    The CPU leg is bounded: it evaluates the first --heatmap-cpu-points rows of X once (after a warm call) and scales
    linearly to the cell's batch (the algorithm is a loop over points; said in every line's cpu_note)."""
    sys.path.insert(0, str(ROOT / "benchmarks"))
    import heatmap
    from oracle import oracle

    cores = oracle.use_all_cores()

    def cpu_timer(layout, X):
        n = min(len(X), args.heatmap_cpu_points)
        xs = np.ascontiguousarray(X[:n])
        oracle.evaluate(layout, xs[: min(n, 8)])
        t0 = time.perf_counter()
        oracle.evaluate(layout, xs)
        dt = time.perf_counter() - t0
        return dt * len(X) / n, f"oracle on {cores} cores, {n} of {len(X)} rows timed, scaled linearly"

    lines, grid = heatmap.run_grid(quick=args.heatmap_quick, seed=0, cpu_timer=cpu_timer, cache_dir=args.heatmap_cache or None,
                                   log=lambda s: print(s, file=sys.stderr, flush=True))
    if args.heatmap_out:
        Path(args.heatmap_out).write_text("".join(json.dumps(l) + "\n" for l in lines))
    heatmap.print_panels(lines, grid)
    heatmap.print_panels(lines, grid, key="log10_cpu_over_b200", title="log10(t_cpu / t_b200)", fmt="{:8.2f}")
    ratios = [l["log10_cpu_over_b200"] for l in lines]
    emit({"metric": "heat-map grid: log10(t_cpu / t_b200) per cell", "cells": len(lines), "min": min(ratios), "median": float(np.median(ratios)),
          "max": max(ratios), "cpu_baseline": {"kind": "port", "cores": cores, "sample": f"first {args.heatmap_cpu_points} rows per cell, scaled"},
          "protocol": "benchmarking/benchmark.py:104-136 (N_ITER = 20, host NumPy in and out)"})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--points", type=int, default=0, help="points per GPU (default: the config's batch)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-standin", action="store_true", help="skip the torch restatement of the reference's algorithm on the GPU")
    ap.add_argument("--no-others", action="store_true", help="skip the records of the other configurations and the cfg5 sweep")
    ap.add_argument("--sweep-points", type=int, default=100_000_000, help="total points of the cfg5 sweep (BASELINE configs[4])")
    ap.add_argument("--standin-points", type=int, default=20_000)
    ap.add_argument("--heatmap", action="store_true", help="run the reference's benchmark grid with a CPU baseline column instead")
    ap.add_argument("--heatmap-quick", action="store_true", help="corners of the grid only")
    ap.add_argument("--heatmap-out", default="", help="JSON lines, one per cell")
    ap.add_argument("--heatmap-cache", default="", help="directory of per-cell .npz timings (reused when present)")
    ap.add_argument("--heatmap-cpu-points", type=int, default=100)
    ap.add_argument("--shard", default="points", choices=["points", "columns"],
                    help="multi-GPU partition: points (weak scaling, tables replicated) or output columns (strong scaling: "
                         "every rank evaluates all points for its slice of the value table; for huge d_out)")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.heatmap:
        return run_heatmap(args)

    import torch
    import torch.distributed as dist

    from smolyax_b200 import _lib, dist as sdist, workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    numa = sdist.bind_to_gpu_numa_node(local)  # before any page-locked allocation (first touch decides where it lives)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = workloads.CONFIGS[args.config]
    n_points = args.points or min(wl.n_points, 1_000_000)
    d_in, d_out = wl.d_in, wl.d_out

    # ---- set-up: rank 0 evaluates f and assembles the tables once, NCCL broadcast, one device handle per rank ----
    ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=d_out, device=local,
                                        batched_f=True)
    columns = args.shard == "columns" and world > 1
    compact = d_out > 256 or columns  # the padded reference layout of cfg3 (29 GB) cannot be assembled: smx_create_compact
    layout = (ip._assemble_compact if compact else ip._assemble)(wl.target(), {})[0] if rank == 0 else None
    col_lo, col_hi = 0, d_out
    if columns:
        col_lo, col_hi = sdist.shard_columns(d_out, rank, world)
        layout = sdist.scatter_columns(layout, d_out, src=0)
        ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=col_hi - col_lo, device=local)
    else:
        layout = sdist.broadcast_layout(layout, src=0)
    ip.set_layout(layout)
    info = ip.device_info()

    # ---- synthetic inputs, resident in HBM ---------------------------------------------------------------------
    gen = torch.Generator(device="cuda").manual_seed(1234 + (0 if columns else rank))  # column shards see the same points
    x = torch.empty((n_points, d_in), dtype=torch.float64, device="cuda")
    if wl.rule == "leja":
        x.uniform_(-1.0, 1.0, generator=gen)
    else:
        x.normal_(0.0, 2.0 ** -0.5, generator=gen)
    y = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        y = ip(x)
    barrier()
    sampler.start()
    launches0 = _lib.lib.smx_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()  # the kernels are launched on torch's current stream (ip.__call__ passes it through the C-ABI)
    for _ in range(args.steps):
        y = ip(x)
    stop.record()
    barrier()
    clocks = sampler.stop()
    launches = int(_lib.lib.smx_launch_count() - launches0)
    timed_kernel = _lib.last_kernel()  # the instantiation the timed steps ran
    elapsed_ms = sdist.max_over_ranks(start.elapsed_time(stop))
    ms_per_step = elapsed_ms / args.steps
    value = (1 if columns else world) * n_points * d_out / (ms_per_step * 1e-3)

    # ---- parity spot check of what was just timed (rank 0): first rows against the CPU oracle ------------------
    parity = None
    if rank == 0:
        check_layout, cols = (layout, None) if not compact else oracle_layout(wl)
        got = y[:64].cpu().numpy()
        if cols is not None:
            keep = (cols >= col_lo) & (cols < col_hi)  # rank 0's column shard
            got = got[:, cols[keep] - col_lo]
            check_layout = {k: (v[:, keep] if k.startswith("F_") else (v[keep] if k == "offset" else v)) for k, v in check_layout.items()}
        parity = value_parity(check_layout, x[:64].cpu().numpy(), got)

    # ---- end to end through the public API from pinned host memory ------------------------------------------------
    x_host = torch.empty((n_points, d_in), dtype=torch.float64, pin_memory=True)
    x_host.copy_(x)
    # the caller's result buffer, page-locked and reused from step to step (`out=`): allocating 0.8 - 8 GB of page-locked
    # memory inside every call (cfg5, cfg3) costs more than the evaluation
    y_host = torch.empty((n_points, col_hi - col_lo), dtype=torch.float64, pin_memory=True)
    ip(x_host[: min(n_points, 65536)], out=y_host[: min(n_points, 65536)])  # warm the staging buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        y_host = ip(x_host, out=y_host)
    torch.cuda.synchronize()
    e2e_s = sdist.max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
    e2e_value = (1 if columns else world) * n_points * d_out / e2e_s
    assert y_host.shape == (n_points, col_hi - col_lo)
    e2e_same = bool(torch.equal(y_host[:4096], y[:4096].cpu()))  # host pipeline against the device-resident call
    # what the link alone does with the same bytes, all ranks at once: one bare cudaMemcpyAsync of x from the same buffer
    barrier()
    t0 = time.perf_counter()
    x.copy_(x_host, non_blocking=True)
    torch.cuda.synchronize()
    bare_s = sdist.max_over_ranks(time.perf_counter() - t0)

    sweep = None
    if not args.no_others and args.config == "cfg2" and not columns:
        del x_host, y_host, x
        torch.cuda.empty_cache()
        x = None
        try:
            sweep = cfg5_sweep(args, world, rank, local, measured_peaks()[0])
        except Exception as exc:
            sweep = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        d_loc = col_hi - col_lo  # outputs this rank's kernel launch produces
        traffic = traffic_source = None
        tfile = ROOT / "profiles" / "traffic.json"
        if tfile.exists():
            try:
                tj = json.loads(tfile.read_text())
                traffic = tj.get(args.config)
                traffic_source = tj.get("source") if traffic is not None else None
            except Exception:
                traffic = None
        roofline = eval_roofline(info, d_in, d_loc, n_points, ms_per_step, peaks, peak_kind, launches // max(args.steps, 1), traffic, traffic_source,
                                 kernel_name=timed_kernel)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if columns else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{args.config}: {wl.rule} d_in={d_in} d_out={d_out} n={wl.n_target} "
                            f"({info['n_summands']} summands, {info['n_terms']} terms), {n_points} points per GPU",
                "points_per_gpu": n_points,
                "parallelism": f"dp{world} (output columns sharded, every rank sees all points)" if columns
                               else f"dp{world} (points sharded, tables replicated)",
                "l2": f"inputs + outputs ({8e-9 * (d_in + d_out) * n_points:.1f} GB per step) exceed L2 (126 MB); no flush needed",
            },
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * d_in * n_points,
                    "d2h_bytes_per_step": 8 * (col_hi - col_lo) * n_points, "ms_per_step": 1e3 * e2e_s, "steps": args.e2e_steps,
                    "result_buffer": "caller's page-locked buffer (out=), reused", "same_bits_as_device_resident_call": e2e_same,
                    "h2d_gbs_aggregate": world * 8e-9 * d_in * n_points / e2e_s,
                    "bare_cudaMemcpyAsync_of_x": {"ms": 1e3 * bare_s, "h2d_gbs_aggregate": world * 8e-9 * d_in * n_points / bare_s,
                                                  "note": "the same page-locked buffer, all ranks at once, no kernel: what the host's PCIe / memory complex delivers"},
                    "numa": numa},
            "gpu_launches": launches,
            "clocks": clocks,
            "parity": parity,
            "parity_max_rel_vs_oracle_first64": parity["max_rel_vs_oracle"] if parity else None,
            "parity_scaled_by_summand_magnitude_first64": parity["scaled_by_summand_magnitude"] if parity else None,
            "library": _lib.lib.smx_build_info().decode(errors="replace"),
        }
        if sweep is not None:
            line["cfg5_sweep"] = sweep
        if world == 1 and not compact and not args.no_gpu_standin and x is not None:
            # the additional number north_star asks for beside the CPU arm: the reference's algorithm and batching on this
            # GPU.  JAX is not installable here, so it is a torch fp64 restatement (benchmarks/reference_gpu_standin.py).
            try:
                from benchmarks import reference_gpu_standin as standin

                n_s = min(n_points, args.standin_points)
                pps, secs, y_s = standin.timed(layout, x[:n_s])
                diff = float((y_s - y[:n_s]).abs().max() / y[:n_s].abs().max())
                line["reference_gpu_standin"] = {
                    "value": pps * d_out, "unit": UNIT, "sample": f"first {n_s} points of the same batch, best of 2 ({secs:.3f} s each)",
                    "kind": "torch fp64 eager restatement of the reference's memory-limited vmap/einsum batches (4 GB limit) on "
                            "this GPU; NOT JAX/XLA (not installable offline), none of this package's kernels",
                    "max_diff_vs_this_arm_rel_to_max": diff}
                del y_s
            except Exception as exc:  # (out of memory on a shared GPU, ...): reported, never fatal for the bench line
                line["reference_gpu_standin"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
            torch.cuda.empty_cache()
        if world == 1 and not args.no_others and args.config == "cfg2":
            del y
            torch.cuda.empty_cache()
            line["others"] = other_records(peaks, peak_kind, local)
        if world == 1 and not args.no_cpu_baseline:
            cpu_layout, cpu_cols = (layout, None) if not compact else oracle_layout(wl)
            v, cores, sample, _ = cpu_reference(wl, cpu_layout, args.cpu_seconds, 1, cpu_cols)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
