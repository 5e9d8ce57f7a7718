#!/bin/bash
# kernel experiments: bench tensor mode under several knobs; prints value + phase shares
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --math tensor --no-cpu-baseline"
run() { echo "== $1"; shift; env "$@" timeout 600 $B 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['value'], l['config']['phase_cycle_share'], l['config']['fast_path_fallbacks'])"; }
run token RG_X=0
run no_token RG_NO_PASS1_TOKEN=1
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or c3_shape or user_views" 2>&1 | tail -3
