#!/bin/bash
# Round-end style check on one B200: full GPU tests, smoke, default bench (+ reference arm), other workloads, ncu launch
# list + full capture of the headline kernel.  usage: scripts/gpu_final.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-final}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/gpu.txt
echo "== pytest gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
echo "== smoke"
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_$TAG.log
echo "== bench default"
timeout -k 10 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default_$TAG.json
echo "== bench reference arm"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference_$TAG.json
echo "== bench c5 sweep"
timeout -k 10 900 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c5_$TAG.json
echo "== bench c2"
timeout -k 10 900 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2_$TAG.json
echo "== bench c4 (200k users)"
timeout -k 10 900 python bench.py --workload c4 --users 200000 --train-users 50 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c4_$TAG.json
echo "== ncu launch list"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rollout -c 12 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --users 200000 --train-users 20 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
tail -1 gpurun_out/ncu_launch_$TAG.log | cut -c1-200
echo "== ncu full capture"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:rollout_cta -s 3 -c 1 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --users 300000 --train-users 20 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/ncu_full_$TAG.log | cut -c1-200
