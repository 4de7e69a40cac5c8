mkdir -p gpurun_out
CMD="python tools/tc_selftest.py --group all --only $1 --batch 32 --precs ${2:-bf16x3}"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 2 -f -o gpurun_out/prof_${3:-layer} $CMD > gpurun_out/ncu_run.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_plain.log; tail -5 gpurun_out/ncu_run.log
