#!/bin/bash
# A/B of an environment knob in one session: usage gpu_ab.sh VAR v1 v2 ... (full 2^20 bench per value, two rounds)
var=$1; shift
for rep in 1 2; do
for v in "$@"; do
  echo -n "$var=$v "
  env $var=$v timeout 600 python bench.py --steps 3 --warmup 2 --cpu-sample 16 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':round(d['value']),'e2e':round(d['e2e']['value']),'ms':round(d['ms_per_step'],2),'phase_ms':{k:round(v,1) for k,v in r['phase_ms'].items()}}))"
done; done
