mkdir -p gpurun_out
CMD="python tools/one_search.py 65536 96 8192"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:nearest_tc -s 1 -c 1 -f -o gpurun_out/prof_search $CMD > gpurun_out/ncu_search.log 2>&1
echo "rc=$?"
