mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu -k "receiver or golden" ; echo "rc=$?" ) > gpurun_out/pytest_rx.log 2>&1
tail -25 gpurun_out/pytest_rx.log | cut -c1-220
