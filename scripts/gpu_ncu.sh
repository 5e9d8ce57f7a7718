#!/bin/bash
# ncu --set full capture of ONE launch of the tensor-mode rollout at the benched size (1M users, C3), optionally with one
# CTA per SM.  usage: scripts/gpu_ncu.sh TAG [users] ; RG_FORCE_ONE_CTA=1 for the single-CTA form
set -u
mkdir -p gpurun_out
TAG=${1:-ncu}
USERS=${2:-1000000}
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:rollout_cta -s 1 -c 1 \
    -o gpurun_out/prof_$TAG -f python scripts/prof_pass1.py --users $USERS > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-300
ls -la gpurun_out/prof_$TAG.ncu-rep
