mkdir -p gpurun_out
( timeout 300 python tools/tc_selftest.py --group all --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_new.log 2>&1
grep -E "FAIL|SELFTEST|rc=|rror" gpurun_out/selftest_new.log | head -3
