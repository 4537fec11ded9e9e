bash scripts/gpu_variants.sh 131072 build/lib_call.so build/lib_call2.so build/lib_call2_m3.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_verify_miller' -c 1 -o gpurun_out/prof_r01_miller_call2 env BN254_B200_LIB=$PWD/build/lib_call2.so python bench.py --n 131072 --steps 1 --warmup 1 --cpu-sample 16 > gpurun_out/ncu_full3.log 2>&1
