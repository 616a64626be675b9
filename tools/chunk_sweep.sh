#!/bin/bash
# Frames per launch (ISX_CHUNK) against throughput: bash tools/chunk_sweep.sh "12 16 20 23 32"
for ch in ${1:-12 16 20 23 32}; do
  for wl in ${WLS:-unary_b64 pairwise_b64}; do
    ISX_CHUNK=$ch python bench.py --steps ${STEPS:-8} --workload $wl --no-cpu-baseline --no-extra | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $ch $wl', round(d['value']), round(d['e2e']['value']), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})"
  done
done
