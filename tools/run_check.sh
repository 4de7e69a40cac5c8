mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu ; echo "rc=$?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ; echo "rc=$?" ) > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
( timeout 600 python bench.py --steps 5 --warmup 3 ; echo "rc=$?" ) > gpurun_out/bench_tc.log 2>&1
cut -c1-220 gpurun_out/bench_tc.log
