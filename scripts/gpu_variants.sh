#!/bin/bash
# usage: gpu_variants.sh <n> <lib...> : short bench of each tuning build, phase times only
n=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib" | tee -a gpurun_out/variants.log
  BN254_B200_LIB=$PWD/$lib timeout 600 python bench.py --n $n --steps 2 --warmup 3 --cpu-sample 16 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':d['value'],'e2e':d['e2e']['value'],'frac':r['frac'],'phase_ms':r['phase_ms'],'clocks':d['clocks']}))" | tee -a gpurun_out/variants.log
done
