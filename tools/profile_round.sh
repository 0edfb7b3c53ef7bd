#!/bin/bash
# One profiling pass on a GPU box (run under gpurun, ONE GPU):  bash tools/profile_round.sh <tag>
#   1. launch list of the bench command (gpu__time_duration per launch; cold-cache + serialised under ncu:
#      the kernels' SHARES of the step are meaningful, not the absolute times)
#   2. ncu --set full of the headline sketch kernel and of the dist kernel inside bench.py (with source)
#   3. ncu --set full of one launch of every sketch / dist kernel variant (tools/bench_configs.py --profile),
#      exported to CSV on the box; the big reports are deleted there (gpurun_out/ is capped at 64 MiB)
# Summarise with tools/ncu_summary.py into profiles/.  Numbers taken under ncu are never bench values.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
B="python bench.py --genomes 400 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-ingest"
$NCU --metrics gpu__time_duration.sum -k regex:'sketch_kernel|dist_|ml_finish|card_|regmin|build_invalid|merge_kernel|nccl' -c 400 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_launches.log 2>&1
$NCU --set full --import-source on -k regex:sketch_kernel --launch-skip 3 -c 1 -f -o $out/${tag}_sketch_ull10_k16 $B > $out/${tag}_ncu1.log 2>&1
$NCU --set full --import-source on -k regex:dist_fgra_tab --launch-skip 3 -c 1 -f -o $out/${tag}_dist_fgra $B > $out/${tag}_ncu2.log 2>&1
$NCU --set full -k regex:'sketch_kernel|build_invalid_mask|dist_kernel|dist_fgra_tab|dist_ml_tab|ml_finish|dist_hll_fast|card_|regmin' -f -o /tmp/${tag}_cfg python tools/bench_configs.py --profile > $out/${tag}_ncu3.log 2>&1
ncu -i /tmp/${tag}_cfg.ncu-rep --page raw --csv > $out/${tag}_cfg_raw.csv 2>> $out/${tag}_ncu3.log
ls -la $out /tmp/${tag}_cfg.ncu-rep
