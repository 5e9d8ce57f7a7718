#!/bin/bash
# tensor-core mode: parity tests for the mode, then a short bench. usage: scripts/gpu_tc.sh TAG
mkdir -p gpurun_out
TAG=${1:-tc}
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or c3_shape or user_views" 2>&1 | tail -5 | tee gpurun_out/tc_test_$TAG.log
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --math tensor --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tensor_$TAG.log
