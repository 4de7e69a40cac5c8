mkdir -p gpurun_out
for U in enc1:bf16x3 dec4:bf16; do
  u=${U%%:*}; p=${U##*:}
  for v in 1 0; do
  ( B2C_RU_W1RES=$v B2C_TC_DEBUG=8 timeout 200 python tools/ru_trace.py --unit $u --prec $p --batch 64 ; echo "rc=$?" ) > gpurun_out/trace_${u}_${p}_w$v.log 2>&1
  head -2 gpurun_out/trace_${u}_${p}_w$v.log
  done
done
