mkdir -p gpurun_out
CMD="python tools/one_search.py 65536 96 8192"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"nearest" -s 4 -c 8 --csv --log-file gpurun_out/search_launches.csv $CMD > gpurun_out/ncu_search.log 2>&1
echo "rc=$?"; grep -E "nearest" gpurun_out/search_launches.csv | cut -d, -f5,13,15 | cut -c1-160
