#!/bin/bash
# Runs on the GPU box under gpurun: parity tests, smoke, a short bench, and an ncu launch list.
# Usage: scripts/gpu_check.sh [bench users per step] [extra pytest args]
set -u
mkdir -p gpurun_out
USERS=${1:-200000}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/gpu.txt
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 3 --warmup 3 --users $USERS 2>&1 | tail -3 | tee gpurun_out/bench.log
