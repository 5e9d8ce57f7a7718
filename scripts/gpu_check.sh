#!/bin/bash
# Runs on the GPU box under gpurun: parity tests, smoke, a short bench, ncu launch list + full capture.
# Usage: scripts/gpu_check.sh [bench users per step] [tag]
set -u
mkdir -p gpurun_out
USERS=${1:-1000000}
TAG=${2:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/gpu.txt
echo "== pytest gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_$TAG.log
echo "== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_$TAG.log
echo "== bench fast"
timeout -k 10 900 python bench.py --steps 3 --warmup 3 --users $USERS --math fast --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_$TAG.log
echo "== bench tensor"
timeout -k 10 900 python bench.py --steps 3 --warmup 3 --users $USERS --math tensor 2>&1 | tail -3 | tee gpurun_out/bench_tensor_$TAG.log
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rollout -c 12 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --users 200000 --train-users 20 --math ${MATH:-fast} \
    --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
tail -3 gpurun_out/ncu_launch_$TAG.log
echo "== ncu full capture of the rollout kernel"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:rollout_cta -s 3 -c 1 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --users 500000 --train-users 20 --math ${MATH:-fast} \
    --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
fi
