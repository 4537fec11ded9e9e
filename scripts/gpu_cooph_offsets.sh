# half-warp layout (two 16-item groups per sub-partition): does a start offset between the two groups of a sub-partition buy overlap?
mkdir -p gpurun_out
for off in 0 3000 7000 10000 14000 18000 21000; do
  BN254_COOP_H=1 BN254_COOP_STAGGER=$off python bench.py --n 262144 --steps 2 --warmup 2 --no-extras --cpu-sample 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cooph offset $off', d['value'], d['roofline']['phase_ms'])"
done
python bench.py --n 262144 --steps 2 --warmup 2 --no-extras --cpu-sample 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('coop4 default', d['value'], d['roofline']['phase_ms'])"
