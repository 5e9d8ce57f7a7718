#!/bin/bash
# parity tests + the C4-shape bench (200k users) + the default bench.  usage: scripts/gpu_c4.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-c4}
timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.log
timeout -k 10 600 python bench.py --workload c4 --users 200000 --train-users 200 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_c4_$TAG.json | cut -c1-100
timeout -k 10 600 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_c3_$TAG.json | cut -c1-100
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
for w in ("c4", "c3"):
    try:
        d = json.loads(open("gpurun_out/bench_%s_%s.json" % (w, tag)).read().strip().splitlines()[-1]); c = d["config"]
        print(w, "%.4g" % d["value"], "mufu", round(d["roofline_mufu"]["frac"], 4), c["phase_cycle_share"], "fallbacks", c.get("fast_path_fallbacks"), c.get("tier3_resamples"))
    except Exception as ex:
        print(w, "failed", ex)
PY
