#!/bin/bash
# pass-1 hand-shake profile (experiment build) + a baseline bench line. usage: scripts/gpu_prof1.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-p1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
export RG_LIB_PATH=$PWD/recogym_b200/librecogym_b200_prof.so
timeout -k 10 300 python scripts/prof_pass1.py --tag c3_2cta 2>&1 | tail -1 > gpurun_out/prof_c3_2cta_$TAG.json
RG_FORCE_ONE_CTA=1 timeout -k 10 300 python scripts/prof_pass1.py --tag c3_1cta 2>&1 | tail -1 > gpurun_out/prof_c3_1cta_$TAG.json
timeout -k 10 300 python scripts/prof_pass1.py --P 100000 --K 64 --users 100000 --tag c4 2>&1 | tail -1 > gpurun_out/prof_c4_$TAG.json
unset RG_LIB_PATH
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_default_$TAG.json
head -c 600 gpurun_out/prof_c3_2cta_$TAG.json; echo
head -c 300 gpurun_out/bench_default_$TAG.json; echo
