#!/bin/bash
# --set full + source-line captures of the per-tile tail kernels (NMS, per-tile merge) over the one-group bench command.
# usage: bash tools/profile_tail.sh <tag>   -> gpurun_out/<tag>_{nms_tiles,merge_tile}_kernel_*.{csv,txt}
set -u
TAG=${1:-tail}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --mosaic 8704 --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-file --no-alt"
$CMD > /dev/null 2>&1 || { echo "bench failed"; exit 1; }
for K in nms_tiles_kernel merge_tile_kernel; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 1 -c 1 \
        -f -o $O/${TAG}_$K $CMD > $O/${TAG}_ncu_$K.log 2>&1
    ncu -i $O/${TAG}_$K.ncu-rep --page details --csv > $O/${TAG}_${K}_details.csv 2>/dev/null
    ncu -i $O/${TAG}_$K.ncu-rep --page source --csv --print-source cuda,sass > $O/${TAG}_${K}_source.csv 2>/dev/null
    python tools/ncu_lines.py $O/${TAG}_${K}_source.csv 40 > $O/${TAG}_${K}_lines.txt 2>&1
    rm -f $O/${TAG}_${K}_source.csv $O/${TAG}_$K.ncu-rep
done
ls -la $O | grep ${TAG}_
