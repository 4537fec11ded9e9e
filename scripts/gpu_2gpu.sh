#!/bin/bash
# 2-GPU pass: sharded host logic on real GPUs, config 5 at 2^22 pairs (strong scaling), bench line at N=2 (weak scaling)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_gpu_check.py 2>&1 | tail -2
timeout 900 $TR --master-port 29512 scripts/bench_distinct.py 22 2>&1 | tail -1 | tee gpurun_out/distinct_2gpu.json | cut -c1-400
timeout 1200 $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_2gpu.json | cut -c1-700
