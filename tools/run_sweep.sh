mkdir -p gpurun_out
( timeout 600 python tools/search_sweep.py --quick --out gpurun_out/search_sweep_quick.json ; echo "rc=$?" ) > gpurun_out/search_sweep_quick.log 2>&1
tail -32 gpurun_out/search_sweep_quick.log | cut -c1-220
