mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu ; echo "rc=$?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
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
ncu --metrics gpu__time_duration.sum --clock-control none -s 810 -c 290 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
