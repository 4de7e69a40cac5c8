mkdir -p gpurun_out
for D in ${DBGS:-0 7}; do
  echo "== B2C_TC_DEBUG=$D"
  B2C_TC_DEBUG=$D timeout 200 python tools/tc_selftest.py --group ru --batch 64 --only "d1" 2>&1 | grep -E "^ru" | cut -c1-60,95-125,170-200
done
