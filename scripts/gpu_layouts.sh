#!/bin/bash
# layouts of the cooperative machine: parity of the pairing modes under each default, then the bench at 2^18 triples
mkdir -p gpurun_out
for env in "BN254_COOP_GROUPS4=0" "BN254_COOP_GROUPS4=1"; do
  echo "== $env"
  env $env timeout 900 python -m pytest tests -m gpu -x -q -k "pairing_modes or verify_batch_with or full_size_verify" 2>&1 | tail -2
  env $env timeout 600 python bench.py --n 262144 --steps 2 --warmup 3 --cpu-sample 16 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':d['value'],'e2e':d['e2e']['value'],'frac':r['frac'],'phase_ms':r['phase_ms'],'clocks':d['clocks']}))"
done
