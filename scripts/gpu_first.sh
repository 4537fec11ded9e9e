#!/bin/bash
# first GPU pass: parity tests, calibration, a short bench, launch list and one full ncu capture of the Miller kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 300 tools/microbench.bin > gpurun_out/microbench.json 2> gpurun_out/microbench.err
cat gpurun_out/microbench.json
( time timeout 900 python bench.py --n 65536 --steps 2 --warmup 3 ) > gpurun_out/bench_64k.log 2>&1
tail -3 gpurun_out/bench_64k.log
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1
tail -2 gpurun_out/bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --n 16384 --steps 1 --warmup 1 --cpu-sample 16 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_verify_miller|k_final_exp_check' -c 2 -o gpurun_out/prof_r01 python bench.py --n 16384 --steps 1 --warmup 1 --cpu-sample 16 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
