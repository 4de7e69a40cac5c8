mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu ; echo "rc=$?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/prof_tc.json ; echo "rc=$?" ) > gpurun_out/bench_tc.log 2>&1
cut -c1-200 gpurun_out/bench_tc.log
python - <<'P'
import json
d=json.load(open('gpurun_out/prof_tc.json'))
by={}
for o in d['launches']: by[o['kind']]=by.get(o['kind'],0)+o['ms']
print({k:round(v,3) for k,v in by.items()}, round(sum(by.values()),3))
P
( timeout 600 python tools/latency.py --reps 300 --out gpurun_out/latency.json ; echo "rc=$?" ) > gpurun_out/latency.log 2>&1
cut -c1-200 gpurun_out/latency.log
( timeout 300 python tools/profile_program.py --batch 1 > gpurun_out/prof_b1.log 2>&1; tail -40 gpurun_out/prof_b1.log )
