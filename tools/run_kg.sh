mkdir -p gpurun_out
for cfg in "0 0" "0 1" "0 2" "0 3" "0 4" "32 1" "32 2" "32 3"; do
  set -- $cfg
  echo "== BK=$1 KGROUP=$2"
  ( [ "$1" != 0 ] && export B2C_RU_BK=$1; [ "$2" != 0 ] && export B2C_RU_KGROUP=$2; timeout 200 python tools/tc_selftest.py --group ru --batch 32 --only "d1" 2>&1 | grep -E "^ru" | sed -E 's/ \| f32[^|]*//; s/err raw [^|]*//g' | cut -c1-160 )
done
