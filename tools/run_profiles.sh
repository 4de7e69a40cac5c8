# launch list (every launch with its device time) + one full ncu capture of the top kernel; bench first without ncu
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 260 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/bench_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 130 -c 3 -f -o gpurun_out/prof_bench_conv $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
