#!/bin/bash
# kernel experiments: bench tensor mode under several builds / knobs; prints value + the raw totals[10:16] shares
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --math tensor --no-cpu-baseline"
run() { echo "== $1"; shift; env "$@" timeout 600 $B 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['value'], l['config']['phase_cycle_share'], l['config']['fast_path_fallbacks'])"; }
run default RG_X=0
run diag RG_LIB_PATH=$PWD/recogym_b200/librecogym_b200_diag.so
