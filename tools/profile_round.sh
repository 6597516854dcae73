#!/bin/bash
# Profiling passes of one round (run on the GPU box through gpurun; every pass only after the plain command exited 0).
# usage: bash tools/profile_round.sh <tag>      -> gpurun_out/<tag>_*.{csv,json,log}
# Recipe: /opt/skills/guides/B200_PROFILING.md (launch list with gpu__time_duration.sum, --set full of the top kernels,
# --clock-control none everywhere; numbers printed by a run under ncu are never bench values).
set -u
TAG=${1:-prof}
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --mosaic 8704 --steps 1 --warmup 1 --no-cpu-baseline --no-profile"
FWD="python tools/profile_layers.py --variant l --batch 296 --reps 1"
$BENCH > $O/${TAG}_bench_8704.json 2> $O/${TAG}_bench_8704.err || { echo "bench failed"; exit 1; }
# 1. launch list of the bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/${TAG}_launches_bench_8704.csv \
    $BENCH > $O/${TAG}_ncu_launches.log 2>&1
# 2. dram bytes + tensor-pipe activity of every conv launch of one forward (stem + 102 conv_igemm launches)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
    --clock-control none -k regex:'conv_igemm_kernel|stem_conv_kernel' -c 103 --csv --log-file $O/${TAG}_conv_metrics_per_launch.csv \
    $FWD > $O/${TAG}_ncu_conv_metrics.log 2>&1
# 3. --set full of representative conv launches (index among conv_igemm launches of a forward):
#    2 = model.2.m.0.cv1 (3x3, 64 -> 64 at 160^2: N = 64 tiles), 26 = model.6.m.0.cv1 (3x3, 256 -> 256 at 40^2: CTA pairs),
#    87 = model.22.cv3.0.0 (3x3, 256 -> 256 at 80^2, detect head)
for IDX in 2 26 87; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel --launch-skip $IDX -c 1 \
        -f -o $O/${TAG}_conv_full_$IDX $FWD > $O/${TAG}_ncu_conv_full_$IDX.log 2>&1
    ncu -i $O/${TAG}_conv_full_$IDX.ncu-rep --page details --csv > $O/${TAG}_conv_full_${IDX}_details.csv 2>/dev/null
    rm -f $O/${TAG}_conv_full_$IDX.ncu-rep      # gpurun brings back at most 64 MiB: keep the CSV exports only
done
# 4. --set full of the preprocessing kernels (first group of the bench command) and the stem
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pp_bucket_kernel|pp_chain_kernel|pp_fused4_kernel' -c 3 \
    -f -o $O/${TAG}_pp_full $BENCH > $O/${TAG}_ncu_pp_full.log 2>&1
ncu -i $O/${TAG}_pp_full.ncu-rep --page details --csv > $O/${TAG}_pp_full_details.csv 2>/dev/null
ncu -i $O/${TAG}_pp_full.ncu-rep --page raw --csv > $O/${TAG}_pp_full_raw.csv 2>/dev/null
rm -f $O/${TAG}_pp_full.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stem_conv_kernel|sppf_pool3_kernel|upsample2_kernel|score_key_kernel|nms_tiles_kernel' -c 6 \
    -f -o $O/${TAG}_misc_full $BENCH > $O/${TAG}_ncu_misc_full.log 2>&1
ncu -i $O/${TAG}_misc_full.ncu-rep --page details --csv > $O/${TAG}_misc_full_details.csv 2>/dev/null
rm -f $O/${TAG}_misc_full.ncu-rep
ls -la $O | grep ${TAG}_
