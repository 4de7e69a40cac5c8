# bench.py on N GPUs of one box, both arms, as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 ; echo "rc=$?" ) > gpurun_out/bench_${N}gpu.log 2>&1
grep -E '^\{|rc=' gpurun_out/bench_${N}gpu.log | cut -c1-330
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 ; echo "rc=$?" ) > gpurun_out/bench_${N}gpu_ref.log 2>&1
grep -E '^\{|rc=' gpurun_out/bench_${N}gpu_ref.log | cut -c1-200
