mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu ; echo "rc=$?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for P in 1 0; do
( B2C_PDL=$P timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ; echo "rc=$?" ) > gpurun_out/bench_pdl$P.log 2>&1
echo "PDL=$P"; cut -c1-200 gpurun_out/bench_pdl$P.log
( B2C_PDL=$P timeout 600 python tools/latency.py --reps 300 --out gpurun_out/latency_pdl$P.json ; echo "rc=$?" ) > gpurun_out/latency_pdl$P.log 2>&1
cut -c1-200 gpurun_out/latency_pdl$P.log | head -2
done
