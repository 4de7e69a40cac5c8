mkdir -p gpurun_out
run() { echo "== $*"; env "$@" > gpurun_out/tmp_bench.log 2>&1; python - <<'P'
import json
for l in open('gpurun_out/tmp_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['e2e']['value'],1), d['ms_per_step'], d['clocks']['sm_mhz'], d['kernel_time_ms_per_program'])
        break
else:
    print(open('gpurun_out/tmp_bench.log').read()[-1500:])
P
}
run A=1 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline
run B2C_LANES=1 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline
run A=1 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline --micro-batch 128
run B2C_LANES=1 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline --micro-batch 128
run A=1 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline --micro-batch 128 --batch 256
