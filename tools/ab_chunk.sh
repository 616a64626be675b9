for ch in 16 32 64; do
  for w in unary_b64 pairwise_b64; do
    echo "== ISX_CHUNK=$ch $w"
    ISX_CHUNK=$ch python bench.py --workload $w --no-extra --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'u16', round(d['e2e_u16']['value']))"
  done
done
python tools/fullsize_parity.py --out gpurun_out/r2_fullsize_parity.json
