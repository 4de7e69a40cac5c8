mkdir -p gpurun_out
( timeout 300 python tools/tc_selftest.py --group all --batch 2 ; echo "rc=$?" ) > gpurun_out/selftest_b2.log 2>&1
grep -E "FAIL|SELFTEST|rc=|rror|timeout" gpurun_out/selftest_b2.log | head
( timeout 300 python tools/tc_selftest.py --group all --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_all.log 2>&1
grep -E "FAIL|SELFTEST|rc=|rror|timeout" gpurun_out/selftest_all.log | head
