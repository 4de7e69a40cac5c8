for v in 1 0; do
  echo "== W7RES=$v"
  B2C_RU_W7RES=$v timeout 300 python tools/tc_selftest.py --group ru --batch 64 --only enc1 --precs bf16x3 2>&1 | cut -c1-175
  B2C_RU_W7RES=$v timeout 200 python tools/power_probe.py --only "enc1" --no-program --secs 2.0 2>&1 | grep -v Warn | tail -1
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_residual or stages_teacher or batch_invariance" 2>&1 | tail -3
