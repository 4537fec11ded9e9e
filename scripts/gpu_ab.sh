#!/bin/bash
# A/B of a runtime knob (BN254_COOP_STAGGER) in one session: alternate the two settings three times
for rep in 1 2 3; do
for st in 0 2147483648; do
  echo -n "knob=$st "
  BN254_COOP_STAGGER=$st timeout 600 python bench.py --n 262144 --steps 3 --warmup 2 --cpu-sample 16 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':round(d['value']),'coop_ms':round(r['phase_ms']['miller_and_final_exp'],2)}))"
done; done
