#!/bin/bash
# re-entry full pass: parity tests, default bench, reference arm, smoke, launch list, full ncu capture of k_verify_lines + k_coop_run
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
( time timeout 1200 python bench.py ) > gpurun_out/bench_full.log 2>&1; tail -4 gpurun_out/bench_full.log | cut -c1-1800
kill $SMI
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; tail -3 gpurun_out/bench_ref.log | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --n 262144 --steps 2 --warmup 1 --cpu-sample 16 > gpurun_out/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_coop4_run|k_verify_lines" -c 2 -o gpurun_out/prof_r01_v8 python bench.py --n 131072 --steps 1 --warmup 1 --cpu-sample 16 > gpurun_out/ncu_prof_v8.log 2>&1
tail -2 gpurun_out/ncu_prof_v8.log | cut -c1-300
