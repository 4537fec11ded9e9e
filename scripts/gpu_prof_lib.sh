#!/bin/bash
# usage: gpu_prof_lib.sh <lib.so> <out name>: full ncu capture of k_coop_run from a tuning build (results may be wrong on purpose)
mkdir -p gpurun_out
BN254_BENCH_NOCHECK=1 BN254_B200_LIB=$PWD/$1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_coop_run" -c 1 -o gpurun_out/$2 python bench.py --n 131072 --steps 1 --warmup 1 --cpu-sample 16 > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log | cut -c1-200
