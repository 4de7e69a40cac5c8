mkdir -p gpurun_out
for v in 1 0; do
  B2C_RU_W1RES=$v timeout 300 python tools/tc_selftest.py --group ru --batch 64 > gpurun_out/ru_w1res$v.log 2>&1
  echo "W1RES=$v"; cut -c1-250 gpurun_out/ru_w1res$v.log
done
