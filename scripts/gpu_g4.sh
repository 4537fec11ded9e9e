#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -4
timeout 1200 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_g4.json | cut -c1-900
timeout 900 python scripts/bench_distinct.py 22 2>&1 | tail -1 | cut -c1-300
