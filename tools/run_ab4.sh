mkdir -p gpurun_out
( timeout 300 python tools/tc_selftest.py --group all --batch 64 --only "block.2.block.1" ; echo "rc=$?" ) > gpurun_out/selftest_a.log 2>&1
( B2C_TC_SLAB=2 B2C_TC_SLAB_WIDE=1 timeout 300 python tools/tc_selftest.py --group all --batch 64 --only "block.2.block.1" ; echo "rc=$?" ) > gpurun_out/selftest_b.log 2>&1
grep -E "FAIL|SELFTEST|rc=|rror|timeout" gpurun_out/selftest_a.log gpurun_out/selftest_b.log | head -8
