#!/bin/bash
# 2-GPU pass: the NCCL paths of bn254_b200/dist.py with the CUDA engine (forged / undecodable items on the last rank), then the bench
# line at N = 2 (weak scaling; its `configs` record carries config 4 and config 5 sharded over the two ranks)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_gpu_worker.py 2>&1 | tail -2
timeout 1200 $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_2gpu.json | cut -c1-700
