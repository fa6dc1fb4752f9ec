#!/bin/bash
M=gpu__time_duration.sum,sm__cycles_elapsed.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tc.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:cov_tf32 -c 12 --csv python tools/probe_cov.py 2>/dev/null | python3 -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]; i=h.index('Metric Name'); v=h.index('Metric Value'); k=h.index('ID'); g=h.index('Grid Size')
seen=set()
for r in rows[1:]:
    if r[g] in seen and r[i]=='gpu__time_duration.sum': pass
    key=(r[g],r[i])
    if key in seen: continue
    seen.add(key)
    print('%-16s %-90s %s'%(r[g],r[i],r[v]))
"
