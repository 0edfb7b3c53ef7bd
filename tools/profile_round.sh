#!/bin/bash
# One profiling pass on a GPU box (run under gpurun, ONE GPU):  bash tools/profile_round.sh <tag>
#   1. launch list of the bench command (gpu__time_duration per launch; cold-cache + serialised under ncu:
#      the kernels' SHARES of the step are meaningful, not the absolute times)
#   2. one launch of every hot kernel at the BASELINE shapes (tools/profile_kernels.py) under ncu with the counters the
#      rooflines need -> gpurun_out/<tag>_kernels_raw.csv + manifest; tools/kernel_costs.py turns them into
#      profiles/kernel_costs.json, tagged with the hash of the CUDA sources (bench.py refuses a stale file)
#   3. ncu --set full (with source) of the headline sketch kernel and the FGRA dist kernel inside bench.py
# Numbers taken under ncu are never bench values.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
B="python bench.py --genomes 400 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-ingest --legs ''"
M="smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,launch__registers_per_thread,launch__grid_size"
eval $NCU --metrics gpu__time_duration.sum -k regex:"'sketch_kernel|dist_|ml_finish|card_|regmin|build_invalid|merge_kernel|text_|out_checksum|nccl'" -c 400 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_launches.log 2>&1
$NCU --metrics $M -k regex:'sketch_kernel|build_invalid_mask|dist_kernel|dist_fgra_tab|dist_ml_tab|ml_finish|dist_hll_fast|dist_hll_int|dist_hmh_fast|text_' --csv --page raw --log-file $out/${tag}_kernels_raw.csv python -m tools.profile_kernels $out/${tag}_manifest.jsonl > $out/${tag}_kernels.log 2>&1
python -m tools.kernel_costs $out/${tag}_manifest.jsonl $out/${tag}_kernels_raw.csv profiles/${tag}_kernels_raw.csv $out/${tag}_kernel_costs.json >> $out/${tag}_kernels.log 2>&1
if [ -z "$SKIP_FULL" ]; then
eval $NCU --set full --import-source on -k regex:sketch_kernel --launch-skip 3 -c 1 -f -o /tmp/${tag}_sketch $B > $out/${tag}_ncu1.log 2>&1
ncu -i /tmp/${tag}_sketch.ncu-rep --page raw --csv > $out/${tag}_sketch_ull10_k16_raw.csv 2>> $out/${tag}_ncu1.log
eval $NCU --set full --import-source on -k regex:dist_fgra_tab --launch-skip 3 -c 1 -f -o /tmp/${tag}_dist $B > $out/${tag}_ncu2.log 2>&1
ncu -i /tmp/${tag}_dist.ncu-rep --page raw --csv > $out/${tag}_dist_fgra_raw.csv 2>> $out/${tag}_ncu2.log
fi
tail -3 $out/${tag}_kernels.log
ls -la $out | tail -12
