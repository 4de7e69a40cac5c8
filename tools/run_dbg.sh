mkdir -p gpurun_out
for D in ${DBGS:-0 7}; do
  echo "== B2C_TC_DEBUG=$D"
  B2C_TC_SLAB=0 B2C_TC_DEBUG=$D timeout 200 python tools/tc_selftest.py --group all --batch 32 --only "$1" 2>&1 | grep -v SELFTEST | cut -c1-50,100-330
done
