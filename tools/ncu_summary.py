#!/usr/bin/env python
"""Key metrics of one `ncu --set full` capture: `ncu -i X.ncu-rep --page raw --csv > raw.csv; ncu_summary.py raw.csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, zip(units, r)))
    for k in KEYS:
        if k in d:
            print(f"{k:<96}{d[k][1]} {d[k][0]}")
    for k in sorted(d):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k][1] or 0) >= 0.05:
            print(f"{k:<96}{d[k][1]}")
