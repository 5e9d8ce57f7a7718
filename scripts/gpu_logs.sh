#!/bin/bash
# GPU: log pipeline tests + micro-benchmark (+ optional ncu launch list of the log kernels). usage: scripts/gpu_logs.sh TAG [ncu]
set -u
mkdir -p gpurun_out
TAG=${1:-logs}
timeout -k 10 900 python -m pytest tests/test_gpu_logs.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_logs_$TAG.log
timeout -k 10 600 python scripts/bench_logs.py --users 1000000 --P 100 --K 5 2>&1 | tail -8 | tee gpurun_out/bench_logs_c2_$TAG.log
timeout -k 10 600 python scripts/bench_logs.py --users 300000 --P 10000 --K 20 2>&1 | tail -8 | tee gpurun_out/bench_logs_c3_$TAG.log
if [ "${2:-}" = "ncu" ]; then
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'scan_|order_scatter|columns_kernel|organic_pairs|bandit_stats|row_argmax|ips_' -c 40 --csv \
    --log-file gpurun_out/launches_logs_$TAG.csv python scripts/bench_logs.py --users 300000 --P 100 --K 5 --reps 1 \
    > gpurun_out/ncu_logs_$TAG.log 2>&1
tail -2 gpurun_out/ncu_logs_$TAG.log | cut -c1-300
fi
