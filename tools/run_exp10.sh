P=multimodal_vqvae_compression_audio_tactile_b200/libb2c.so
for nap in 0 32 128; do
  cp gpurun_libs/libb2c_nap$nap.so $P
  echo "== nap $nap"
  timeout 200 python tools/power_probe.py --only "enc1,enc2,k7 C256,k1 C256,dec4" --no-program --secs 1.5 2>&1 | grep -v Warn | tail -5
  python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
done
cp gpurun_libs/libb2c_nap0.so $P
echo "== nap0 DEBUG=7"
B2C_TC_DEBUG=7 timeout 200 python tools/power_probe.py --only "enc1" --no-program --secs 1.5 2>&1 | grep -v Warn | tail -1
cp gpurun_libs/libb2c_nap128.so $P
echo "== nap128 DEBUG=7"
B2C_TC_DEBUG=7 timeout 200 python tools/power_probe.py --only "enc1" --no-program --secs 1.5 2>&1 | grep -v Warn | tail -1
