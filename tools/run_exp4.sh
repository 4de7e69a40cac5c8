mkdir -p gpurun_out
for w in 1 0; do for g in 1 2 3; do
  echo "== W1RES=$w KGROUP=$g"
  B2C_RU_W1RES=$w B2C_RU_KGROUP=$g timeout 300 python tools/tc_selftest.py --group ru --batch 64 --only enc1 --precs bf16x3 2>&1 | cut -c1-175
  B2C_RU_W1RES=$w B2C_RU_KGROUP=$g timeout 300 python tools/tc_selftest.py --group ru --batch 64 --only dec4.d1 --precs bf16 2>&1 | cut -c1-175
done; done
echo "== STACK=0 W1RES=1"
for g in 1 2; do B2C_RU_STACK=0 B2C_RU_W1RES=1 B2C_RU_KGROUP=$g timeout 300 python tools/tc_selftest.py --group ru --batch 64 --only enc1 --precs bf16x3 2>&1 | cut -c1-175; done
echo "== SLAB=0 W1RES=1"
for g in 1 2; do B2C_RU_SLAB=0 B2C_RU_W1RES=1 B2C_RU_KGROUP=$g timeout 300 python tools/tc_selftest.py --group ru --batch 64 --only enc1 --precs bf16x3 2>&1 | cut -c1-175; done
