timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_variants.sh 131072 bn254_b200/libbn254_b200.so build/lib_ilp_m3.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_verify_miller' -c 1 -o gpurun_out/prof_r01_miller_ilp python bench.py --n 131072 --steps 1 --warmup 1 --cpu-sample 16 > gpurun_out/ncu_full5.log 2>&1
