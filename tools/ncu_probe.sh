#!/bin/bash
# ncu pipe / stall summary of ONE dist kernel on a tools/dist_probe.py case (run under gpurun, one GPU):
#   bash tools/ncu_probe.sh <kernel regex> <dist_probe args...>     e.g.  bash tools/ncu_probe.sh dist_ml_tab ull-ml 4000
# Numbers taken under ncu are never bench values.
M="smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,launch__registers_per_thread,smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio,smsp__average_warp_latency_issue_stalled_not_selected.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio,smsp__average_warp_latency_issue_stalled_mio_throttle.ratio"
K=${1:-dist_hll_int}
shift
mkdir -p gpurun_out
ncu --clock-control none --metrics $M -k regex:"$K" --csv --page raw --log-file gpurun_out/probe_ncu.csv python tools/dist_probe.py "$@" --profile > gpurun_out/probe_ncu.log 2>&1
tail -1 gpurun_out/probe_ncu.log
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/probe_ncu.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
names = rows[hdr]
for r in rows[hdr + 2:]:
    print({n.replace('.avg.pct_of_peak_sustained_active', '%').replace('.avg.pct_of_peak_sustained_elapsed', '%el').replace('smsp__average_warp_latency_issue_stalled_', 'stall_').replace('.ratio', ''): v
           for n, v in zip(names, r) if n == 'Kernel Name' or ('__' in n and not n.startswith(('device__', 'launch__occ', 'profiler__', 'nvlink__', 'numa__', 'c2clink__')) and
                                                                  ('.avg.pct' in n or 'stalled' in n or n.endswith('.sum') or n == 'launch__registers_per_thread'))})
PY
