timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash scripts/gpu_variants.sh 1048576 bn254_b200/libbn254_b200.so build/lib_lines3.so build/lib_lines4.so
