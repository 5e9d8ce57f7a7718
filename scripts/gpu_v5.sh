bash scripts/gpu_try.sh v6a
bash scripts/gpu_variants.sh x7 1000000 base kb3
