timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1200 python scripts/bench_configs.py 1048576 2>&1 | tee gpurun_out/bench_configs.jsonl | grep '"5:' | cut -c1-300
