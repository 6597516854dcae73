set -u
O=gpurun_out; TAG=r02ak
FWD="python tools/profile_layers.py --variant l --batch 296 --reps 1"
$FWD > $O/${TAG}_layers_v8l_b296.txt 2>&1 || { echo "layers failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
    --clock-control none -k regex:'conv_igemm_kernel|stem_conv_kernel' -c 103 --csv --log-file $O/${TAG}_conv_metrics_per_launch.csv $FWD > $O/${TAG}_ncu_conv_metrics.log 2>&1
for IDX in 2 87; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel --launch-skip $IDX -c 1 -f -o $O/${TAG}_conv_full_$IDX $FWD > $O/${TAG}_ncu_conv_full_$IDX.log 2>&1
    ncu -i $O/${TAG}_conv_full_$IDX.ncu-rep --page details --csv > $O/${TAG}_conv_full_${IDX}_details.csv 2>/dev/null
    rm -f $O/${TAG}_conv_full_$IDX.ncu-rep
done
ls $O | grep $TAG
