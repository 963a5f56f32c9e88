#!/usr/bin/env python
"""Small readers for Nsight Compute exports (used to write the summaries under profiles/).

    ncu -i X.ncu-rep --page raw --csv    > raw.csv      ->  python profiles/ncu_tools.py raw raw.csv
    ncu -i X.ncu-rep --page source --csv > source.csv   ->  python profiles/ncu_tools.py source source.csv [N]
"""
import csv
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
    "l1tex__data_pipe_lsu_wavefronts_mem_lgds", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread ", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active", "smsp__average_warps_issue_stalled_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier", "smsp__average_warps_issue_stalled_math_pipe", "smsp__average_warps_issue_stalled_wait",
    "smsp__average_warps_issue_stalled_mio_throttle", "smsp__average_warps_issue_stalled_lg_throttle",
    "smsp__average_warps_issue_stalled_not_selected", "smsp__inst_executed.sum ", "sm__cycles_elapsed.max ",
]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        for h, u, v in zip(hdr, units, vals):
            if any((h + " ").startswith(k) or k.strip() == h for k in RAW_KEYS):
                print(f"  {h} [{u}] = {v}")


def source(path, top=30):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    ci = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[start + 1:] if len(r) > ci["# Samples"] and r[ci["# Samples"]] not in ("", "# Samples")]
    key = ci["# Samples"]
    tot = sum(float(r[key]) for r in body) or 1.0
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(f"total samples {tot:.0f}")
    for r in sorted(body, key=lambda r: -float(r[key]))[:top]:
        stalls = sorted(((float(r[ci[s]] or 0), s) for s in stall_cols), reverse=True)[:2]
        st = " ".join(f"{s[6:]}={v:.0f}" for v, s in stalls if v > 0)
        print(f"{float(r[key]) / tot * 100:5.1f}%  {r[ci['Source']][:90]:90s} {st}")


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2])
    else:
        source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
