#!/usr/bin/env python
"""Benchmark of the hot path: batched evaluation of the barycentric Smolyak interpolant.

    python bench.py --gpus N --steps K --warmup W            # this framework (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm (CPU oracle port) on host cores

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): Leja nodes, d_in = 1000, d_out = 1,
n = 10^4 nodes, 10^6 evaluation points per GPU (weak scaling), synthetic inputs U(-1,1)^d, fp64.
A step = one `__call__` over the whole batch = ONE fused kernel launch.  `value` is measured with the batch resident
in HBM (8 GB of x per step, far larger than the 126 MB L2, so no flush is needed); `e2e` is the same batch
evaluated through the public API from pinned HOST memory, copies in and out inside the timed region.
Prints one JSON line (rank 0).
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
    """Reference layout (the oracle's input) of the workload; for huge d_out restricted to ORACLE_COLUMNS evenly spaced
    output columns (the padded F_n of all 10^4 outputs of cfg3 would be 29 GB).  Returns (layout, columns or None)."""
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from smolyax_b200 import workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    wl = workloads.CONFIGS[args.config]
    layout, cols = oracle_layout(wl)
    for _ in range(min(args.warmup, 1)):
        cpu_reference(wl, layout, 0.5, 1, cols)
    value, cores, sample, times = cpu_reference(wl, layout, args.cpu_seconds, max(1, min(args.steps, 3)), cols)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * min(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: {wl.rule} d_in={wl.d_in} d_out={wl.d_out} n={wl.n_target}", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference algorithm (padded per-summand second-barycentric-form contraction) restated in C + OpenMP "
                "(oracle/smx_oracle.c); the reference's JAX runtime is not installable offline",
    }
    emit(line)


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
    ap.add_argument("--standin-points", type=int, default=20_000)
    ap.add_argument("--shard", default="points", choices=["points", "columns"],
                    help="multi-GPU partition: points (weak scaling, tables replicated) or output columns (strong scaling: "
                         "every rank evaluates all points for its slice of the value table; for huge d_out)")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from smolyax_b200 import _lib, dist as sdist, workloads
    from smolyax_b200.interpolation import SmolyakBarycentricInterpolator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
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
    elapsed_ms = sdist.max_over_ranks(start.elapsed_time(stop))
    ms_per_step = elapsed_ms / args.steps
    value = (1 if columns else world) * n_points * d_out / (ms_per_step * 1e-3)

    # ---- parity spot check of what was just timed (rank 0): first rows against the CPU oracle ------------------
    parity = parity_scaled = None
    if rank == 0:
        from oracle import oracle

        check_layout, cols = (layout, None) if not compact else oracle_layout(wl)
        xs = x[:64].cpu().numpy()
        ref = oracle.evaluate(check_layout, xs)
        got = y[:64].cpu().numpy()
        if cols is not None:
            keep = (cols >= col_lo) & (cols < col_hi)  # rank 0's column shard
            got, ref = got[:, cols[keep] - col_lo], ref[:, keep]
        parity = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
        # the same difference in the norm the 1e-12 bound is stated in (tests/test_gpu_parity.py): relative to the summand
        # magnitude sum_nu |zeta_nu| I_nu f (the oracle run with |zeta|; f > 0 for this target family) - the sum has
        # sum |zeta| ~ 2e4 times more rounding noise than its value suggests
        mag_layout = {k: (np.abs(v) if k.startswith("zetas_") or k == "offset" else v) for k, v in check_layout.items()}
        mag = oracle.evaluate(mag_layout, xs)
        if cols is not None:
            mag = mag[:, keep]
        parity_scaled = float(np.max(np.abs(got - ref) / np.maximum(np.abs(mag), 1e-300)))

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

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        d_loc = col_hi - col_lo  # outputs this rank's kernel launch produces
        alg_bytes = 8.0 * (d_in + d_loc) * n_points  # bytes_eval = 8 (d_in + d_out) per point
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tfile = ROOT / "profiles" / "traffic.json"
        if tfile.exists():
            try:
                traffic = json.loads(tfile.read_text()).get(args.config)
            except Exception:
                traffic = None
        dense = bool(info["has_dense_path"])
        fma_per_eval = info["dense_terms"] if dense else info["padded_fma"]
        fp64_tflops = 2.0 * fma_per_eval * d_loc * n_points / (ms_per_step * 1e-3) / 1e12
        fp64_peak = None
        try:
            fp64_peak = json.loads((ROOT / "profiles" / "fp64_peaks.json").read_text())["fp64_dmma_tflops"]
        except Exception:
            pass
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
            "kernel": "fast_multi_kernel" if 2 <= d_loc <= 6 else "fast_lean_kernel", "algorithmic_bytes_per_launch": alg_bytes,
            "fp64_tflops_executed": fp64_tflops, "fp64_dmma_peak_tflops_measured": fp64_peak,
            "note": "x is streamed once (8*(d_in+d_out) B per point); the FP64 tensor work (2*padded_fma flop per point) "
                    f"needs {2.0 * fma_per_eval * d_loc * n_points / ((fp64_peak or 37.12) * 1e12) * 1e3:.2f} ms per launch at the measured DMMA peak, "
                    f"the x stream {alg_bytes / (peaks['hbm_gbs'] * 1e9) * 1e3:.2f} ms at the measured HBM peak",
        }
        # which roof: SURVEY 8(d) t_roof = max(bytes / BW, flops / FP64 rate), with the folded form's 2 * n_terms * d_out
        # flops per point (the fewest any polynomial evaluation of this interpolant needs; padding is not counted)
        fp64_bound = (not dense and fp64_peak and
                      2.0 * info["n_terms"] * d_loc * n_points / (fp64_peak * 1e12) > alg_bytes / (peaks["hbm_gbs"] * 1e9))
        if dense or fp64_bound:
            # GEMM regime (folded form): algorithmic flops = 2 * d_out * n_terms per point, on the FP64 tensor
            # instruction; the denominator is the FP64 DMMA rate measured on this GPU type (profiles/fp64_peaks.json) -
            # MEASURED_PEAKS.json only carries the bf16 tensor rate, which no fp64 path can use.
            alg_flops = 2.0 * info["n_terms"] * d_loc * n_points
            tf = alg_flops / (ms_per_step * 1e-3) / 1e12
            roofline = {
                "bound": "tensor", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak if fp64_peak else None,
                "traffic": traffic, "peak_source": "FP64 DMMA (mma.sync.m8n8k4.f64) rate measured with profiles/fp64_peaks.cu",
                "kernel": "dense_eval_kernel" if dense else roofline["kernel"], "algorithmic_flops_per_launch": alg_flops,
                "hbm_gbs_algorithmic": achieved, "hbm_peak_gbs": peaks["hbm_gbs"],
            }
            if not dense:
                roofline["fp64_tflops_executed"] = fp64_tflops  # block-sparse form incl. its padding (2 * padded_fma per point and output)
                roofline["launches_per_step"] = launches // max(args.steps, 1)
                if ms_per_step < 0.2:
                    roofline["note"] = "a step of this size is launch/latency-bound (tens of microseconds per call), not pipe-bound"
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
                    "result_buffer": "caller's page-locked buffer (out=), reused", "same_bits_as_device_resident_call": e2e_same},
            "gpu_launches": launches,
            "clocks": clocks,
            "parity_max_rel_vs_oracle_first64": parity,
            "parity_scaled_by_summand_magnitude_first64": parity_scaled,
        }
        if world == 1 and not compact and not args.no_gpu_standin:
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
        if world == 1 and not args.no_cpu_baseline:
            cpu_layout, cpu_cols = (layout, None) if not compact else oracle_layout(wl)
            v, cores, sample, _ = cpu_reference(wl, cpu_layout, args.cpu_seconds, 1, cpu_cols)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
