mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu -k "nearest" ; echo "rc=$?" ) > gpurun_out/pytest_nearest.log 2>&1
tail -3 gpurun_out/pytest_nearest.log
( timeout 900 python tools/search_sweep.py --out gpurun_out/search_sweep.json ; echo "rc=$?" ) > gpurun_out/search_sweep.log 2>&1
grep -E "N= 65536|N= 16384" gpurun_out/search_sweep.log | grep -E "D= 96|D=128|D= 64|D= 32" | cut -c1-200
tail -2 gpurun_out/search_sweep.log
