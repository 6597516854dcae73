#!/bin/bash
# --set full + source-page captures of the preprocessing kernels (one launch each) over tools/pp_bench.py.
# usage: bash tools/profile_pp.sh <tag>   -> gpurun_out/<tag>_pp_*.csv
set -u
TAG=${1:-pp}
O=gpurun_out
mkdir -p $O
CMD="python tools/pp_bench.py --fp16 --reps 1"
$CMD > $O/${TAG}_pp_bench.json 2> $O/${TAG}_pp_bench.err || { echo "pp_bench failed"; exit 1; }
for K in pp_bucket_kernel pp_chain_kernel pp_fused4_kernel; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 2 -c 1 \
        -f -o $O/${TAG}_$K $CMD > $O/${TAG}_ncu_$K.log 2>&1
    ncu -i $O/${TAG}_$K.ncu-rep --page details --csv > $O/${TAG}_${K}_details.csv 2>/dev/null
    ncu -i $O/${TAG}_$K.ncu-rep --page raw --csv > $O/${TAG}_${K}_raw.csv 2>/dev/null
    ncu -i $O/${TAG}_$K.ncu-rep --page source --csv --print-source cuda,sass > $O/${TAG}_${K}_source.csv 2>/dev/null
    python tools/ncu_lines.py $O/${TAG}_${K}_source.csv 60 > $O/${TAG}_${K}_lines.txt 2>&1
    gzip -f $O/${TAG}_${K}_source.csv
    rm -f $O/${TAG}_$K.ncu-rep
done
ls -la $O | grep ${TAG}_
