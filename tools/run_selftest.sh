mkdir -p gpurun_out
( timeout 200 python tools/tc_selftest.py --group pred ; echo "rc=$?" ) > gpurun_out/selftest_pred.log 2>&1
( timeout 400 python tools/tc_selftest.py --group enc ; echo "rc=$?" ) > gpurun_out/selftest_enc.log 2>&1
( timeout 400 python tools/tc_selftest.py --group dec ; echo "rc=$?" ) > gpurun_out/selftest_dec.log 2>&1
tail -5 gpurun_out/selftest_pred.log gpurun_out/selftest_enc.log gpurun_out/selftest_dec.log
