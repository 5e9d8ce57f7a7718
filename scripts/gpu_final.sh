#!/bin/bash
# Round-end style check on one B200: full GPU tests, smoke, default bench (+ reference arm), other workloads, the ncu
# launch list of the DEFAULT bench command (1M users per launch: lpt_* + rollout shares of the step are visible) and one
# --set full capture of the 1M-user rollout launch (-s 2: the log-generation and 250k-user warm-up launches are skipped) (+ the same with the L2-persisting window).  usage: scripts/gpu_final.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-final}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
nproc >> gpurun_out/gpu_$TAG.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/gpu_$TAG.txt
if [ -z "${SKIP_TESTS:-}" ]; then
echo "== pytest gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.log
fi
echo "== smoke"
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_$TAG.log
echo "== bench default"
timeout -k 10 900 python bench.py 2>/dev/null | tail -1 | tee gpurun_out/bench_default_$TAG.json | cut -c1-300
echo "== bench reference arm"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_reference_$TAG.json | cut -c1-300
echo "== bench c5 / c2 / c4 (1 GPU)"
timeout -k 10 900 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_c5_$TAG.json | cut -c1-200
timeout -k 10 900 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_c2_$TAG.json | cut -c1-200
timeout -k 10 900 python bench.py --workload c4 --users 200000 --train-users 200 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_c4_$TAG.json | cut -c1-200
echo "== log pipeline (C2 and C3 shapes) incl. generate_logs -> DataFrame end to end"
timeout -k 10 600 python scripts/bench_logs.py --users 1000000 --P 100 --K 5 2>/dev/null > gpurun_out/bench_logs_c2_$TAG.jsonl; tail -1 gpurun_out/bench_logs_c2_$TAG.jsonl | cut -c1-300
timeout -k 10 600 python scripts/bench_logs.py --users 200000 --P 10000 --K 20 --df-users 50000 2>/dev/null > gpurun_out/bench_logs_c3_$TAG.jsonl; tail -1 gpurun_out/bench_logs_c3_$TAG.jsonl | cut -c1-300
echo "== ncu launch list of the default bench command"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
tail -1 gpurun_out/ncu_launch_$TAG.log | cut -c1-200
echo "== ncu full capture (1M users, two CTAs per SM)"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:rollout_cta -s 2 -c 1 \
    -o gpurun_out/prof_$TAG -f python scripts/prof_pass1.py --trained > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/ncu_full_$TAG.log | cut -c1-200
echo "== dram bytes with the L2-persisting window"
RG_L2_PERSIST=1 timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:rollout_cta -s 2 -c 1 --csv --log-file gpurun_out/l2persist_$TAG.csv python scripts/prof_pass1.py --trained > /dev/null 2>&1
tail -2 gpurun_out/l2persist_$TAG.csv | cut -c1-300
RG_L2_PERSIST=1 timeout -k 10 300 python scripts/prof_pass1.py --trained 2>/dev/null | tail -1 | cut -c1-200
