mkdir -p gpurun_out
( timeout 1200 python tools/search_sweep.py --out gpurun_out/search_sweep.json ; echo "rc=$?" ) > gpurun_out/search_sweep.log 2>&1
tail -4 gpurun_out/search_sweep.log | cut -c1-200
