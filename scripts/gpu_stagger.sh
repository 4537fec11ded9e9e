#!/bin/bash
# start-offset sweep (ns per warp index) for the warps of a group in k_coop4_run
for st in 0 50 100 200 400 800; do
  echo "== stagger $st"
  BN254_COOP_STAGGER=$st timeout 600 python bench.py --n 262144 --steps 2 --warmup 2 --cpu-sample 16 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':d['value'],'frac':r['frac'],'coop_ms':r['phase_ms']['miller_and_final_exp']}))"
done
