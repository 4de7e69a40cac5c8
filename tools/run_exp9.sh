timeout 200 python tools/power_probe.py --only "enc1,enc2,k1 C256" --no-program --secs 2.0 2>&1 | grep -v Warn | tail -4
timeout 300 python tools/tc_selftest.py --group ru --batch 8 --precs bf16x3 --only enc 2>&1 | cut -c1-180
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu ; echo "rc=$?" ) 2>&1 | tail -5
timeout 300 python tools/diag_parity.py tc cal_b8k512 2>&1 | tail -12
