#!/bin/bash
# 8-GPU pass: the NCCL paths at world size 8, then the bench line at N = 8 (weak scaling; config 5 = 2^22 pairs over 8 ranks in `configs`)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 tests/dist_gpu_worker.py 2>&1 | tail -2
timeout 900 $TR --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_8gpu.json | cut -c1-400
