mkdir -p gpurun_out
for cfg in "SLAB=0 DBG=0" "SLAB=1 DBG=0" "SLAB=0 DBG=2" "SLAB=0 DBG=1" "SLAB=0 DBG=4"; do
  set -- $cfg
  echo "== $1 $2"
  ( export B2C_RU_${1}; export B2C_TC_DEBUG=${2#DBG=}; timeout 200 python tools/tc_selftest.py --group ru --batch 32 --only "d1" 2>&1 | grep -E "^ru" | sed -E 's/ \| f32[^|]*//; s/err raw ([^ ]*) act [^ |]*/err \1/g' | cut -c1-170 )
done
