#!/bin/bash
# usage: tools/ncu_summary.sh <tag>   (reads gpurun_out/<tag>_prof.ncu-rep; writes raw/src csv beside it and prints a summary)
T=$1; cd "$(dirname "$0")/../gpurun_out"
ncu -i ${T}_prof.ncu-rep --page raw --csv > ${T}_raw.csv 2>/dev/null
ncu -i ${T}_prof.ncu-rep --page source --csv --print-source sass,cuda > ${T}_src.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('${T}_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active']
for h,u,v in zip(hdr,units,vals):
    if h in want: print(f"{h:95s} {v} {u}")
PY
