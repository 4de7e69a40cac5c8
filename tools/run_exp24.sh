for v in 2 1; do
  echo "== STGLOCK=$v"
  B2C_RU_STGLOCK=$v timeout 300 python tools/tc_selftest.py --group ru --batch 64 --only enc2 --precs bf16x3 2>&1 | cut -c1-175
  B2C_RU_STGLOCK=$v timeout 200 python tools/power_probe.py --only "enc2" --no-program --secs 2.0 2>&1 | grep -v Warn | tail -1
done
