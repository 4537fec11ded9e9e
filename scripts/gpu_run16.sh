timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash scripts/gpu_variants.sh 262144 bn254_b200/libbn254_b200.so
