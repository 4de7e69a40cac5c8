mkdir -p gpurun_out
( timeout 300 python tools/diag_parity.py tc c3_b10k128 ; echo "rc=$?" ) > gpurun_out/diag_tc_c3.log 2>&1
( timeout 300 python tools/diag_parity.py tc cal_b8k512 ; echo "rc=$?" ) > gpurun_out/diag_tc_cal.log 2>&1
( timeout 600 python bench.py --precision tc --steps 3 --warmup 3 --profile-out gpurun_out/prof_tc.json ; echo "rc=$?" ) > gpurun_out/bench_tc.log 2>&1
