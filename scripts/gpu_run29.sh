#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rlc or distinct or aggregate" 2>&1 | tail -12
timeout 1200 python scripts/bench_configs.py 1048576 2>&1 | tee gpurun_out/bench_configs.jsonl | grep -E "random|5:|Error|error" | cut -c1-300
