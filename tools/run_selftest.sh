mkdir -p gpurun_out
B=${1:-32}
( timeout 300 python tools/tc_selftest.py --group all --batch $B ; echo "rc=$?" ) > gpurun_out/selftest_all.log 2>&1
grep -c "TF/s" gpurun_out/selftest_all.log; grep -E "FAIL|SELFTEST|rc=|rror" gpurun_out/selftest_all.log | head
