#!/bin/bash
# The two BASELINE configurations that are DEFINED on eight GPUs, launched as the driver launches bench.py:
#   configs[3]: P = 100 000, K = 64, 10M users over 8 GPUs (1.25M per GPU per step), bandit_count agent (sparse statistics)
#   configs[4]: A/B sweep of the 5 leaderboard agents x 1M users (125k per GPU per agent), P = 10 000, K = 20
# plus the default C3 line at N GPUs.  usage: scripts/gpu_8.sh N TAG
N=${1:-8}; TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/gpus_$TAG.txt
run() {  # name, port, args...
  local name=$1 port=$2; shift 2
  timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N "$@" 2>gpurun_out/err_${name}_$TAG.log | tail -1 | tee gpurun_out/bench_${name}_${N}gpu_$TAG.json | cut -c1-400
}
run c4 29531 --workload c4 --steps 2 --warmup 3 --train-users 1000
run c5 29532 --workload c5 --users $((1000000 / N)) --steps 3 --warmup 3
run c3 29533 --steps 5 --warmup 3 --no-cpu-baseline
