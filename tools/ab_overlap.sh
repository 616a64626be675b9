for cfg in "ISX_OVERLAP_TABLES=0" "ISX_OVERLAP_TABLES=1 ISX_TABLES_PRIO=0" "ISX_OVERLAP_TABLES=1 ISX_TABLES_PRIO=1"; do
  for w in unary_b64 pairwise_b64; do
    echo "== $cfg $w"
    env $cfg python bench.py --workload $w --no-extra --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'u16', round(d['e2e_u16']['value']), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()})"
  done
done
