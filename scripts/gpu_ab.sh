# A/B of tuning builds: scripts/gpu_ab.sh name1 name2 ...  (variants/lib_<name>.so), twice each, n = 2^18
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  BN254_B200_LIB=$PWD/variants/lib_$v.so python bench.py --n ${N:-262144} --steps 2 --warmup 2 --no-extras --cpu-sample 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d['roofline']['phase_ms']
print('$v rep$rep value %.0f hash %.2f lines %.2f machine %.2f' % (d['value'], p['hash_to_g1'], p['line_sets'], p['miller_and_final_exp']))"
done
done
