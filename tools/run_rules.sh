mkdir -p gpurun_out
for cfg in "X=0" "B2C_TC_SLAB=2" "B2C_TC_KGROUP=1" "B2C_TC_KGROUP=2" "B2C_TC_KGROUP=4" "B2C_TC_EPI2=0" "B2C_TC_EPI2=1"; do
  echo "== $cfg"
  ( export $cfg; timeout 300 python tools/tc_selftest.py --group all --batch 32 2>&1 | grep -E "^(enc|dec|ru)" | sed -E 's/ \| f32[^|]*//; s/err raw ([^ ]*) act ([^ |]*)//g; s/TF\/s//g; s/ \| \|.*//' | cut -c1-150 )
done
