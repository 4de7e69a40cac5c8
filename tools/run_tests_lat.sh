mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu ; echo "rc=$?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/prof_tc.json ; echo "rc=$?" ) > gpurun_out/bench_tc.log 2>&1
cut -c1-260 gpurun_out/bench_tc.log
( timeout 600 python tools/latency.py --reps 300 --out gpurun_out/latency.json ; echo "rc=$?" ) > gpurun_out/latency.log 2>&1
cut -c1-330 gpurun_out/latency.log
python tools/profile_program.py --batch 1 --top 6 2>&1 | tail -9
