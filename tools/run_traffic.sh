mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain3.log 2>&1 &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_tc_kernel|conv_ru_kernel|conv_tc2_kernel" -s 492 -c 82 --csv --log-file gpurun_out/traffic_conv.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_traffic.log; wc -l gpurun_out/traffic_conv.csv
