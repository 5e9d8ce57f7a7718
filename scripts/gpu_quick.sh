#!/bin/bash
# Quick GPU check: selected parity tests + one tensor-mode bench. usage: scripts/gpu_quick.sh TAG [pytest -k expr]
set -u
mkdir -p gpurun_out
TAG=${1:-q}
KEXPR=${2:-}
if [ -n "$KEXPR" ]; then
  timeout -k 10 1200 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_$TAG.log
else
  timeout -k 10 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_$TAG.log
fi
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --math tensor --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tensor_$TAG.log
