mkdir -p gpurun_out
for U in enc1:bf16x3 dec4:bf16 enc2:bf16x3 dec3:bf16; do
  u=${U%%:*}; p=${U##*:}
  ( B2C_TC_DEBUG=8 timeout 200 python tools/ru_trace.py --unit $u --prec $p --batch 32 ; echo "rc=$?" ) > gpurun_out/trace_${u}_${p}.log 2>&1
  head -3 gpurun_out/trace_${u}_${p}.log
done
