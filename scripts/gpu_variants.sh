#!/bin/bash
# Times experiment variants of the fast rollout (recogym_b200/librecogym_b200_<V>.so, `make variant`) with prof_pass1.py.
# usage: scripts/gpu_variants.sh TAG users V1 V2 ...   ("base" = the product library)
set -u
mkdir -p gpurun_out
TAG=$1; USERS=$2; shift 2
for V in "$@"; do
  if [ "$V" = base ]; then unset RG_LIB_PATH; else export RG_LIB_PATH=$PWD/recogym_b200/librecogym_b200_$V.so; fi
  for ONE in 0 1; do
    RG_FORCE_ONE_CTA=$ONE timeout -k 5 200 python scripts/prof_pass1.py --users $USERS --tag ${V}_one$ONE 2>/dev/null | tail -1 > gpurun_out/var_${TAG}_${V}_$ONE.json
  done
done
python - "$TAG" "$@" <<'PY'
import json, sys
tag = sys.argv[1]
for v in sys.argv[2:]:
    for one in (0, 1):
        try:
            d = json.load(open("gpurun_out/var_%s_%s_%d.json" % (tag, v, one)))
            e = d.get("epilogue_per_tile_per_warp", {}); i = d.get("issuer_per_unit", {})
            print("%-10s one_cta=%d ms %8.2f steps/s %.3e pass1/128draws %9.0f share %s epi(wait %s work %s) iss(tma %s dempty %s issue %s) sub %s" % (
                v, one, d["ms"], d["user_steps_per_s"], d["pass1_cycles_per_128_draws"], d["phase_share"]["pass1"],
                round(e.get("wait_dfull", 0)), round(e.get("tile_total_after_dfull", 0)), round(i.get("wait_tma", 0)),
                round(i.get("wait_dempty", 0)), round(i.get("issue", 0)), d.get("thread0_cycles_per_round", "")))
        except Exception as ex:
            print(v, one, "failed", ex)
PY
