mkdir -p gpurun_out
( timeout 300 python tools/tc_selftest.py --group all --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_all.log 2>&1
grep -E "FAIL|SELFTEST|rc=|rror|timeout" gpurun_out/selftest_all.log | head
( B2C_TC_SLAB=0 timeout 300 python tools/tc_selftest.py --group all --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_all_v1.log 2>&1
grep -E "FAIL|SELFTEST|rc=|rror|timeout" gpurun_out/selftest_all_v1.log | head
