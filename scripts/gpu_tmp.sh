B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
run() { echo "== $1"; shift; env "$@" timeout 600 $B 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['value'], l['config']['pass1_cycles_per_128_draws'], l['config']['phase_cycle_share'], l['config']['fast_path_fallbacks'])"; }
run one_cta_6stages RG_FORCE_ONE_CTA=1
run two_ctas RG_X=1
