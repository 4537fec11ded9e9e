#!/bin/bash
# start-offset sweep for co-resident blocks of k_coop_run (cycles)
for st in 0 800 1600 2400 4000 8000 20000; do
  echo "== stagger $st"
  BN254_COOP_STAGGER=$st timeout 600 python bench.py --n 262144 --steps 2 --warmup 2 --cpu-sample 16 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':d['value'],'frac':r['frac'],'coop_ms':r['phase_ms']['miller_and_final_exp']}))"
done
