#!/bin/bash
# A/B runs of differently tuned builds of the library (build/*.so, see the ISX_* macros in csrc/dp.cu).
for v in "" "$@"; do
  echo "== variant: ${v:-default}"
  for wl in ${WLS:-unary_b64 pairwise_b64}; do
    ISX_LIB_PATH=$v python bench.py --steps 10 --workload $wl --no-cpu-baseline --no-extra | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', round(d['value']), round(d['e2e']['value']), round(d['stage_ms_per_step']['dp'],2))"
  done
done
