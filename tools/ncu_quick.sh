#!/bin/bash
# usage: tools/ncu_quick.sh <out-prefix> [bench args...]   — a light ncu pass (selected metrics, one launch) over the tcgen05 prune kernel
OUT=$1; shift
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_membar_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_sleeping_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct,smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct,smsp__inst_executed.sum,sm__warps_active.avg.per_cycle_active,launch__registers_per_thread
timeout 170 ncu --metrics $M --clock-control none -k regex:${KREGEX:-k_prune_tc5} -c 1 --csv --log-file $OUT.csv python bench.py --no-cpu-baseline --steps 1 --warmup 3 --cols 2097152 "$@" > /dev/null 2>$OUT.err
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print("%-85s %s %s"%(d["Metric Name"],d["Metric Value"],d["Metric Unit"]))
PY
