#!/bin/bash
# N-GPU bench exactly as the driver launches it. usage: scripts/gpu_multi.sh N TAG [noref]
N=${1:-2}; TAG=${2:-mg}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$TAG.txt
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_${N}gpu_$TAG.log
if [ "${3:-}" != "noref" ]; then
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_${N}gpu_$TAG.log
fi
