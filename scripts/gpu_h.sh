#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "pairing_modes" 2>&1 | tail -2
for cfg in "BN254_COOP_H=0" "BN254_COOP_H=1 BN254_COOP_STAGGER=0" "BN254_COOP_H=1 BN254_COOP_STAGGER=6000" "BN254_COOP_H=1 BN254_COOP_STAGGER=12000" "BN254_COOP_H=1 BN254_COOP_STAGGER=18000"; do
  echo "== $cfg"
  env $cfg timeout 600 python bench.py --n 262144 --steps 2 --warmup 2 --cpu-sample 16 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':d['value'],'frac':r['frac'],'coop_ms':r['phase_ms']['miller_and_final_exp']}))"
done
