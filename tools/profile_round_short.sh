set -u
O=gpurun_out; TAG=ai; mkdir -p $O
BENCH="python bench.py --mosaic 8704 --steps 1 --warmup 1 --no-cpu-baseline --no-profile"
FWD="python tools/profile_layers.py --variant l --batch 296 --reps 1"
$BENCH > $O/${TAG}_bench_8704.json 2> $O/${TAG}_bench_8704.err || { echo "bench failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/${TAG}_launches_bench_8704.csv $BENCH > $O/${TAG}_ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'conv_igemm_kernel|stem_conv_kernel' -c 103 --csv --log-file $O/${TAG}_conv_metrics_per_launch.csv $FWD > $O/${TAG}_ncu_conv_metrics.log 2>&1
ls -la $O | grep ${TAG}_
