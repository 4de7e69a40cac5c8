mkdir -p gpurun_out
( timeout 300 python tools/tc_selftest.py --group all --batch 32 ; echo "rc=$?" ) > gpurun_out/selftest_all.log 2>&1
grep -E "FAIL|SELFTEST|rc=|rror" gpurun_out/selftest_all.log | head
P=${1:-tc}
( timeout 600 python bench.py --precision $P --steps 3 --warmup 3 --profile-out gpurun_out/prof_$P.json ; echo "rc=$?" ) > gpurun_out/bench_$P.log 2>&1
cut -c1-300 gpurun_out/bench_$P.log
