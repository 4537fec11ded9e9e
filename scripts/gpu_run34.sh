#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
timeout 1200 python scripts/bench_configs.py 1048576 2>&1 | tee gpurun_out/bench_configs.jsonl | cut -c1-260
