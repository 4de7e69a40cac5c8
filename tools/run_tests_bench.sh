mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu ; echo "rc=$?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/prof_tc.json ; echo "rc=$?" ) > gpurun_out/bench_tc.log 2>&1
cut -c1-300 gpurun_out/bench_tc.log
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ; echo "rc=$?" ) > gpurun_out/bench_ref.log 2>&1
cut -c1-300 gpurun_out/bench_ref.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ; echo "rc=$?" ) > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
