#!/bin/bash
# Guarded development iteration on one B200: smoke first (short timeout: a hung kernel must not eat the GPU budget), then
# the pass-1 profile (experiment build), the GPU tests and one bench line.  usage: scripts/gpu_try.sh TAG [notests]
set -u
mkdir -p gpurun_out
TAG=${1:-try}
MODE=${2:-all}
export RG_DEBUG_LAUNCH=1
timeout -k 5 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke_$TAG.log
if ! grep -q "smoke ok" gpurun_out/smoke_$TAG.log; then echo "SMOKE FAILED"; exit 1; fi
if [ -f recogym_b200/librecogym_b200_prof.so ]; then
  export RG_LIB_PATH=$PWD/recogym_b200/librecogym_b200_prof.so
  timeout -k 5 200 python scripts/prof_pass1.py --tag c3_2cta 2>gpurun_out/prof_err_$TAG.log | tail -1 > gpurun_out/prof_c3_2cta_$TAG.json
  RG_FORCE_ONE_CTA=1 timeout -k 5 200 python scripts/prof_pass1.py --tag c3_1cta 2>/dev/null | tail -1 > gpurun_out/prof_c3_1cta_$TAG.json
  timeout -k 5 200 python scripts/prof_pass1.py --P 100000 --K 64 --users 100000 --tag c4 2>/dev/null | tail -1 > gpurun_out/prof_c4_$TAG.json
  unset RG_LIB_PATH
  grep "\[rg\]" gpurun_out/prof_err_$TAG.log | head -3
fi
python - <<PY
import json
for n in ("prof_c3_2cta","prof_c3_1cta","prof_c4"):
    try:
        d=json.load(open("gpurun_out/%s_$TAG.json"%n)); d.pop("trace",None); d.pop("trace_cols",None); print(json.dumps(d))
    except Exception as e: print(n, "failed", e)
PY
if [ "$MODE" != "notests" ]; then
  timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$TAG.log
fi
timeout -k 10 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_default_$TAG.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_default_$TAG.json")); print("bench value %.4g e2e %.4g ms %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"])); print(d["config"]["phase_cycle_share"], d["config"].get("pass1_cycles_per_128_draws"), d.get("roofline_mufu",{}).get("frac"))
except Exception as e: print("bench failed", e)
PY
