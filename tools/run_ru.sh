mkdir -p gpurun_out
( timeout 300 python tools/tc_selftest.py --group ru --batch 2 --T 4800 ; echo "rc=$?" ) > gpurun_out/selftest_ru_small.log 2>&1
cat gpurun_out/selftest_ru_small.log | cut -c1-260
( timeout 300 python tools/tc_selftest.py --group ru --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_ru.log 2>&1
cat gpurun_out/selftest_ru.log | cut -c1-260
