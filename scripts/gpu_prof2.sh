#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-p2}
export RG_LIB_PATH=$PWD/recogym_b200/librecogym_b200_prof.so
timeout -k 10 300 python scripts/prof_pass1.py --tag c3_2cta 2>&1 | tail -1 > gpurun_out/prof_c3_2cta_$TAG.json
RG_FORCE_ONE_CTA=1 timeout -k 10 300 python scripts/prof_pass1.py --tag c3_1cta 2>&1 | tail -1 > gpurun_out/prof_c3_1cta_$TAG.json
unset RG_LIB_PATH
timeout -k 10 300 python scripts/dbg_host_vs_device.py 2>&1 | tail -4 | tee gpurun_out/host_vs_device_$TAG.txt
python - <<PY
import json
for n in ("prof_c3_2cta","prof_c3_1cta"):
    try:
        d=json.load(open("gpurun_out/%s_$TAG.json"%n)); d.pop("trace",None); d.pop("trace_cols",None); print(json.dumps(d))
    except Exception as e: print(n, "failed", e)
PY
