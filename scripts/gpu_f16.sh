#!/bin/bash
# fp16 vs tf32 operand elements (RG_TC_F16): logit error, parity tests, C3 / C4 bench lines of both builds on the same box.
# usage: scripts/gpu_f16.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-f16}
echo "== tc logits error"
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k tensor_core_logits 2>&1 | grep -E "tc logits|passed|failed|Error|error" | head -20
echo "== pytest gpu"
timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$TAG.log
for V in f16 tf32; do
  if [ "$V" = f16 ]; then unset RG_LIB_PATH; else export RG_LIB_PATH=$PWD/recogym_b200/librecogym_b200_tf32.so; fi
  echo "== $V: bench c3, c4 (200k users), c5"
  timeout -k 10 600 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_c3_${V}_$TAG.json | cut -c1-120
  RG_DEBUG_LAUNCH=1 timeout -k 10 600 python bench.py --workload c4 --users 200000 --train-users 200 --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/err_c4_${V}_$TAG.log | tail -1 | tee gpurun_out/bench_c4_${V}_$TAG.json | cut -c1-120
  grep -m2 "rollout_cta" gpurun_out/err_c4_${V}_$TAG.log
  timeout -k 10 600 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_c5_${V}_$TAG.json | cut -c1-120
done
unset RG_LIB_PATH
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
for w in ("c3", "c4", "c5"):
    for v in ("f16", "tf32"):
        try:
            d = json.loads(open("gpurun_out/bench_%s_%s_%s.json" % (w, v, tag)).read().strip().splitlines()[-1]); c = d["config"]
            print(w, v, "%.4g" % d["value"], "mufu", round(d["roofline_mufu"]["frac"], 4), c["phase_cycle_share"], "fallbacks", c.get("fast_path_fallbacks"), c.get("tier3_resamples"))
        except Exception as ex:
            print(w, v, "failed", ex)
PY
