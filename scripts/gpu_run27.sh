#!/bin/bash
mkdir -p gpurun_out
BN254_COOP_W=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_coopw_run" -c 1 -o gpurun_out/prof_r01_w1 python bench.py --n 124320 --steps 1 --warmup 1 --cpu-sample 16 > gpurun_out/ncu_prof_w1.log 2>&1
tail -2 gpurun_out/ncu_prof_w1.log | cut -c1-300
