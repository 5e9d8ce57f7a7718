#!/bin/bash
# ncu full capture of the fast rollout kernel only. usage: scripts/gpu_prof.sh TAG [users]
set -u
mkdir -p gpurun_out
TAG=${1:-prof}
USERS=${2:-100000}
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:rollout_cta -s 3 -c 1 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --users $USERS --train-users 20 \
    --math ${MATH:-tensor} --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-400
