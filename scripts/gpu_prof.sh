# usage: gpu_prof.sh <kernel regex> <out name> [n] [launch-skip]: one ncu --set full capture of a kernel inside a short bench run, with the
# raw-metric and per-instruction source pages exported as CSV next to the report
n=${3:-131072}
skip=${4:-2}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$1" -s $skip -c 1 -o gpurun_out/$2 python bench.py --n $n --steps 1 --warmup 1 --no-extras --cpu-sample 16 > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log | cut -c1-300
ncu -i gpurun_out/$2.ncu-rep --page raw --csv > gpurun_out/raw_$2.csv 2>/dev/null
ncu -i gpurun_out/$2.ncu-rep --page source --csv > gpurun_out/src_$2.csv 2>/dev/null
ls -la gpurun_out/$2.ncu-rep gpurun_out/raw_$2.csv gpurun_out/src_$2.csv
