mkdir -p gpurun_out
for cfg in ${CFGS:-"128 64" "128 128" "256 128" "192 96"}; do
  set -- $cfg
  ( timeout 600 python bench.py --steps 4 --warmup 3 --batch $1 --micro-batch $2 --no-cpu-baseline ; echo "rc=$?" ) > gpurun_out/bench_mb_$1_$2.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_mb_$1_$2.log').readline())
    print('batch $1 mb $2:', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['kernel_time_ms_per_program'])
except Exception as e:
    print('batch $1 mb $2 failed', e, open('gpurun_out/bench_mb_$1_$2.log').read()[-300:])
PY
done
