mkdir -p gpurun_out
( B2C_RU_SLAB=0 timeout 300 python tools/tc_selftest.py --group ru --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_a.log 2>&1
( timeout 300 python tools/tc_selftest.py --group ru --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_b.log 2>&1
( timeout 300 python tools/tc_selftest.py --group ru --batch 3 --T 5000 ; echo "rc=$?" ) > gpurun_out/selftest_c.log 2>&1
grep -E "FAIL|SELFTEST|rc=|rror|timeout" gpurun_out/selftest_a.log gpurun_out/selftest_b.log gpurun_out/selftest_c.log | head -8
