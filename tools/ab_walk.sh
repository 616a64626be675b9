for v in "" build/lib_walk20.so build/lib_walk24.so; do
  echo "== lib=$v"
  ISX_LIB_PATH=$v python bench.py --workload pairwise_b64 --no-extra --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
ISX_LIB_PATH=build/lib_walk20.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pairwise_tile_walk" 2>&1 | tail -2
