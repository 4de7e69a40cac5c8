mkdir -p gpurun_out
CMD="$1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c ${3:-400} --csv --log-file gpurun_out/${2:-launches}.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain.log
