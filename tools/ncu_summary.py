#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof_r1_bwd.ncu-rep  > profiles/r1_bwd_full.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ,
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit() and r[-3] == "gpu__time_duration.sum"]
    agg = OrderedDict()
    for r in rows:
        name, ns = r[4].split("(")[0].replace("void ", ""), float(r[-1].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot/1e6:.3f} ms of kernel time (ncu: cold-cache, serialised -> compare shares)")
    print(f"{'kernel':70s} {'launches':>8s} {'avg us':>10s} {'share %':>8s}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {n:8d} {ns/n/1e3:10.2f} {100*ns/tot:8.2f}")


def full(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==", r[idx["Kernel Name"]], "grid", r[idx.get("Grid Size", 0)] if "Grid Size" in idx else "")
        for k in KEYS:
            if k in idx:
                print(f"  {k:85s} {r[idx[k]]:>16s} {units[idx[k]]}")
        rd = float(r[idx["dram__bytes_read.sum"]]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[idx["dram__bytes_write.sum"]]]
        print(f"  => DRAM traffic per launch: {(rd+wr)/1e6:.1f} MB")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
