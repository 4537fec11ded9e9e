#!/bin/bash
# warp-local layout of the cooperative machine: parity of the four pairing modes, then block vs warp-local bench at 2^18 triples
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pairing_modes or verify" 2>&1 | tail -5
for w in 0 1; do
  echo "== BN254_COOP_W=$w"
  BN254_COOP_W=$w timeout 600 python bench.py --n 262144 --steps 2 --warmup 3 --cpu-sample 16 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':d['value'],'e2e':d['e2e']['value'],'frac':r['frac'],'phase_ms':r['phase_ms'],'clocks':d['clocks']}))"
done
