#!/usr/bin/env python
"""Key counters of an .ncu-rep (raw page) as a markdown table: ncu_raw_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
name_i = h.index("Kernel Name") if "Kernel Name" in h else None
print("| metric | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " | unit |")
print("|---|" + "---|" * (len(rows) - 1))
for w in WANT:
    if w in h:
        i = h.index(w)
        print(f"| {w} | " + " | ".join(r[i] for r in rows[2:]) + f" | {rows[1][i]} |")
if name_i is not None:
    print("\nkernel:", rows[2][name_i][:100])
