#!/usr/bin/env python
"""Summarise ncu outputs for profiles/ (run here, no GPU needed).
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep   > profiles/<name>_full.txt
  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/<name>_launches.txt
"""
import csv, io, subprocess, sys, collections

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "sass__inst_executed_register_spilling", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads",
    "sass__inst_executed_global_stores",
]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {path} (per launch; cold-cache, serialised - compare shares, not absolutes)")
    for r in rows[2:]:
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:88s} {r[i]} {units[i]}")
        rd = float(r[hdr.index("dram__bytes_read.sum")]) if "dram__bytes_read.sum" in hdr else 0
        wr = float(r[hdr.index("dram__bytes_write.sum")]) if "dram__bytes_write.sum" in hdr else 0
        u = units[hdr.index("dram__bytes_read.sum")]
        print(f"{'traffic = dram read + write':88s} {rd + wr:.6f} {u}")
        print("-" * 120)


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[4].split("(")[0][:90]
        tot[name][0] += 1
        tot[name][1] += float(r[-1])
    total = sum(v[1] for v in tot.values())
    print(f"# launch list {path}: {len(rows)} launches, {total/1e6:.3f} ms of device time (gpu__time_duration.sum)")
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns/1e6:10.3f} ms  {100*ns/total:5.1f}%  x{n:<4d} avg {ns/n/1e3:9.1f} us  {name}")


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2])
