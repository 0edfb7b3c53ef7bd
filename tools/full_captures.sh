#!/bin/bash
# `ncu --set full` of the headline sketch kernel and of the tile kernels changed in round 2, summarised by tools/ncu_summary.py
# (run under gpurun, ONE GPU):  bash tools/full_captures.sh r02     -> gpurun_out/<tag>_full_<kernel>.json
# Numbers taken under ncu are never bench values.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
NCU="ncu --set full --import-source on --clock-control none -c 1 -f"
$NCU -k regex:sketch_kernel --launch-skip 3 -o /tmp/${tag}_sk python bench.py --genomes 400 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-ingest --legs '' > $out/${tag}_full.log 2>&1
python -m tools.ncu_summary /tmp/${tag}_sk.ncu-rep $out/${tag}_full_sketch_ull10_k16.json 1999994000 kmer >> $out/${tag}_full.log 2>&1
$NCU -k regex:dist_fgra_tab -o /tmp/${tag}_fg python tools/dist_probe.py ull-fgra 8000 --profile >> $out/${tag}_full.log 2>&1
python -m tools.ncu_summary /tmp/${tag}_fg.ncu-rep $out/${tag}_full_dist_fgra_tab.json 32772096000 register_pair >> $out/${tag}_full.log 2>&1
$NCU -k regex:dist_ml_tab -o /tmp/${tag}_ml python tools/dist_probe.py ull-ml 8000 --profile >> $out/${tag}_full.log 2>&1
python -m tools.ncu_summary /tmp/${tag}_ml.ncu-rep $out/${tag}_full_dist_ml_tab_gsum.json 32772096000 register_pair >> $out/${tag}_full.log 2>&1
$NCU -k regex:dist_hll_int -o /tmp/${tag}_hl python tools/dist_probe.py hll 4000 --profile >> $out/${tag}_full.log 2>&1
python -m tools.ncu_summary /tmp/${tag}_hl.ncu-rep $out/${tag}_full_dist_hll_int.json 131104768000 register_pair >> $out/${tag}_full.log 2>&1
$NCU -k regex:hmh_ec_gemm -o /tmp/${tag}_ec python tools/dist_probe.py hmh-small 4000 --profile >> $out/${tag}_full.log 2>&1
python -m tools.ncu_summary /tmp/${tag}_ec.ncu-rep $out/${tag}_full_hmh_ec_gemm.json 8002000 small_pair >> $out/${tag}_full.log 2>&1
grep -c . $out/${tag}_full.log; ls -la $out | grep ${tag}_full
