#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "codecs or bn256 or reference_api" 2>&1 | tail -3
timeout 1200 python scripts/bench_configs.py 1048576 2>&1 | tee gpurun_out/bench_configs.jsonl | cut -c1-250
