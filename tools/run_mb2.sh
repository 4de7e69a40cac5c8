mkdir -p gpurun_out
for cfg in "128 64" "192 96" "256 128" "128 128" "96 48"; do
  set -- $cfg
  ( timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --batch $1 --micro-batch $2 ) > gpurun_out/bench_mb_$1_$2.log 2>&1
  python - "$1" "$2" <<'P'
import json,sys
b,m=sys.argv[1],sys.argv[2]
for line in open(f"gpurun_out/bench_mb_{b}_{m}.log"):
    if line.startswith("{"):
        d=json.loads(line); print("batch",b,"micro",m,"value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms/step",round(d["ms_per_step"],2))
P
done
