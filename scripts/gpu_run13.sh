export BN254_BENCH_NOCHECK=1
bash scripts/gpu_variants.sh 262144 build/lib_abl_nobar.so build/lib_abl_noxi.so build/lib_abl_both.so
