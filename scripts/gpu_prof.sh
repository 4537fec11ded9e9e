# usage: gpu_prof.sh <kernel regex> <out name> [n]
n=${3:-131072}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$1" -c 1 -o gpurun_out/$2 python bench.py --n $n --steps 1 --warmup 1 --cpu-sample 16 > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log | cut -c1-300
