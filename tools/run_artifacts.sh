# round-1 artifact run (one B200): tests, bench (both arms), latency, smoke, ncu launch list, ncu DRAM traffic of the
# conv launches of one program, ncu --set full of the fused residual unit
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu ; echo "rc=$?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/prof_tc.json ; echo "rc=$?" ) > gpurun_out/bench_tc.log 2>&1
cut -c1-260 gpurun_out/bench_tc.log
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ; echo "rc=$?" ) > gpurun_out/bench_ref.log 2>&1
cut -c1-200 gpurun_out/bench_ref.log
( timeout 600 python tools/latency.py --reps 500 --out gpurun_out/latency.json ; echo "rc=$?" ) > gpurun_out/latency.log 2>&1
cut -c1-330 gpurun_out/latency.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ; echo "rc=$?" ) > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 810 -c 290 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain3.log 2>&1 &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_tc_kernel|conv_ru_kernel|conv_tc2_kernel" -s 492 -c 82 --csv --log-file gpurun_out/traffic_conv.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"; wc -l gpurun_out/traffic_conv.csv
( timeout 900 python tools/search_sweep.py --out gpurun_out/search_sweep.json ; echo "rc=$?" ) > gpurun_out/search_sweep.log 2>&1
tail -3 gpurun_out/search_sweep.log | cut -c1-300
CMD="python tools/tc_selftest.py --group ru --only enc1.d1 --batch 32 --precs bf16x3"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ru -s 1 -c 1 -f -o gpurun_out/prof_ru_enc1_x3 $CMD > gpurun_out/ncu_run.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_run.log
for U in enc1:bf16x3 enc2:bf16x3 dec3:bf16 dec4:bf16; do
  u=${U%%:*}; p=${U##*:}
  ( B2C_TC_DEBUG=8 timeout 200 python tools/ru_trace.py --unit $u --prec $p --batch 32 ; echo "rc=$?" ) > gpurun_out/trace_${u}_${p}.log 2>&1
  head -1 gpurun_out/trace_${u}_${p}.log
done
