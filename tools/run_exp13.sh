for cfg in "B2C_RU_W7RES=0" "A=1" "B2C_RU_W7RES=0 B2C_RU_DIRECT=0"; do
  echo "== $cfg"
  env $cfg timeout 300 python tools/tc_selftest.py --group ru --batch 64 --only enc1 --precs bf16x3 2>&1 | cut -c1-175
  env $cfg timeout 300 python tools/tc_selftest.py --group ru --batch 64 --only dec4 --precs bf16 2>&1 | cut -c1-175
  env $cfg timeout 200 python tools/power_probe.py --only "enc1,dec4" --no-program --secs 2.0 2>&1 | grep -v Warn | tail -2
done
