mkdir -p gpurun_out
for S in 1 0; do
echo "== B2C_RU_STACK=$S"
( B2C_RU_STACK=$S timeout 300 python tools/tc_selftest.py --group ru --only enc1 --batch 32 2>&1 | grep -E "^ru|SELFTEST" | sed -E 's/ \| f32[^|]*//' | cut -c1-190 )
done
( timeout 1200 python -m pytest tests -x -q -m gpu ; echo "rc=$?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_tc.json ; echo "rc=$?" ) > gpurun_out/bench_tc.log 2>&1
cut -c1-200 gpurun_out/bench_tc.log
( timeout 300 python tools/diag_parity.py > gpurun_out/diag_tc.log 2>&1; tail -12 gpurun_out/diag_tc.log | cut -c1-200 )
