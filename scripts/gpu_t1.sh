#!/bin/bash
( timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
timeout 600 python bench.py --n 262144 --steps 2 --warmup 3 --cpu-sample 16 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'value':d['value'],'e2e':d['e2e']['value'],'frac':r['frac'],'phase_ms':r['phase_ms']}))"
