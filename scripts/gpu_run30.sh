#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_distinct.py 22 2>&1 | tail -2 | tee gpurun_out/distinct_1gpu.json | cut -c1-400
