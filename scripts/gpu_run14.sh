timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_variants.sh 1048576 bn254_b200/libbn254_b200.so
