mkdir -p gpurun_out
( B2C_TC_SLAB=0 timeout 300 python tools/tc_selftest.py --group all --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_g1.log 2>&1
( timeout 300 python tools/tc_selftest.py --group all --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_g2.log 2>&1
( B2C_TC_SLAB=0 B2C_TC_KGROUP=1 timeout 300 python tools/tc_selftest.py --group all --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_g3.log 2>&1
grep -E "FAIL|SELFTEST|rc=|rror" gpurun_out/selftest_g?.log | head
