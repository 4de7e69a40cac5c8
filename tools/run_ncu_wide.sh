mkdir -p gpurun_out
i=0
for spec in "all enc.block.4.block.0.block.1 bf16x3 conv_tc" "all dec.model.1.block.2.block.1 bf16 conv_tc" "ru enc2.d1 bf16x3 conv_ru" "ru dec4.d1 bf16 conv_ru"; do
  set -- $spec
  i=$((i+1))
  CMD="python tools/tc_selftest.py --group $1 --only $2 --batch 32 --precs $3"
  $CMD > gpurun_out/ncu_plain_$i.log 2>&1 &&
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$4 -s 1 -c 1 -f -o gpurun_out/prof_full_$i $CMD > gpurun_out/ncu_run_$i.log 2>&1
  echo "$spec rc=$?"; grep -E "^(enc|dec|ru)" gpurun_out/ncu_plain_$i.log | cut -c1-150
done
