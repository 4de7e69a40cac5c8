mkdir -p gpurun_out
( timeout 300 python tools/tc_selftest.py --group ru --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_ru.log 2>&1
sed -E 's/ \| f32[^|]*//; s/err raw ([^ ]*) act ([^ |]*)/err \1 \2/g' gpurun_out/selftest_ru.log | cut -c1-200
( timeout 600 python tools/tc_selftest.py --group all --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_all.log 2>&1
grep -v "^ru" gpurun_out/selftest_all.log | grep -E "FAIL|SELFTEST|rc=|dec.model.[34]" | sed -E 's/err raw ([^ ]*) act ([^ |]*)/err \1 \2/g' | cut -c1-220
( timeout 600 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/prof_tc.json ; echo "rc=$?" ) > gpurun_out/bench_tc.log 2>&1
cut -c1-400 gpurun_out/bench_tc.log
( timeout 1200 python -m pytest tests -x -q -m gpu ; echo "rc=$?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
