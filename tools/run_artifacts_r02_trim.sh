# round-2 artifact run (one B200): bench (both arms, per-launch profile), ncu launch list of the bench command, ncu DRAM
# traffic of the conv launches, ncu --set full of the round-2 kernels (fused unit with resident weights, tcgen05
# residual VQ, full-length attention of the PLC forward, a wide bf16x3 conv), pipeline trace, power probe
mkdir -p gpurun_out
( timeout 900 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/prof_tc.json ; echo "rc=$?" ) > gpurun_out/bench_tc.log 2>&1
cut -c1-200 gpurun_out/bench_tc.log
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ; echo "rc=$?" ) > gpurun_out/bench_ref.log 2>&1
cut -c1-200 gpurun_out/bench_ref.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/launches_bench.csv
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_tc_kernel|conv_ru_kernel|conv_tc2_kernel" -c 500 --csv --log-file gpurun_out/traffic_conv.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"; wc -l gpurun_out/traffic_conv.csv
CMD="python tools/tc_selftest.py --group ru --only enc1.d1 --batch 32 --precs bf16x3"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ru -s 1 -c 1 -f -o gpurun_out/r02_full_ru_enc1_x3 $CMD > gpurun_out/ncu_run1.log 2>&1
echo "full ru rc=$?"
CMD="python tools/one_rvq.py 64 75 8 512 96"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rvq_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02_full_rvq_tc $CMD > gpurun_out/ncu_run2.log 2>&1
echo "full rvq rc=$?"
for U in enc1:bf16x3; do
  u=${U%%:*}; p=${U##*:}
  ( B2C_TC_DEBUG=8 timeout 200 python tools/ru_trace.py --unit $u --prec $p --batch 64 ; echo "rc=$?" ) > gpurun_out/trace_${u}_${p}.log 2>&1
  head -1 gpurun_out/trace_${u}_${p}.log
done
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
