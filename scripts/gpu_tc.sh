#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k tensor_core 2>&1 | tail -15 | tee gpurun_out/tc_test.log
