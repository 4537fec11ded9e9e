timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_verify_lines|k_hash_round' -c 3 -o gpurun_out/prof_r01_lines_hash python bench.py --n 131072 --steps 1 --warmup 0 --cpu-sample 16 > gpurun_out/ncu_lines.log 2>&1
tail -2 gpurun_out/ncu_lines.log | cut -c1-200
