mkdir -p gpurun_out
P=${1:-tc}
( timeout 600 python bench.py --precision $P --steps 3 --warmup 3 --profile-out gpurun_out/prof_$P.json ; echo "rc=$?" ) > gpurun_out/bench_$P.log 2>&1
cut -c1-400 gpurun_out/bench_$P.log
