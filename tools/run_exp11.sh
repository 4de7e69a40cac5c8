mkdir -p gpurun_out
run() { echo -n "$* : "; env "$@" python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['kernel_time_ms_per_program']; print(round(d['value'],1), d['clocks']['sm_mhz'], 'x3', k['conv_tc_x3'], 'bf16', k['conv_tc'])"
}
run A=1
run B2C_RU_SLAB=1
run B2C_TC_SLAB=2
run A=1
run B2C_TC_SLAB=0
run B2C_RU_SLAB=0
run B2C_TC_EPI2=0
run B2C_RU_W1RES=0
run A=1
