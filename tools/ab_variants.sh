for v in "" build/lib_u4c4.so build/lib_u2c4.so; do
  echo "== variant: ${v:-default}"
  ISX_LIB_PATH=$v python bench.py --steps 10 --no-cpu-baseline --no-extra | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unary', round(d['value']), round(d['e2e']['value']), round(d['stage_ms_per_step']['dp'],2))"
  ISX_LIB_PATH=$v python bench.py --steps 10 --workload pairwise_b64 --no-cpu-baseline --no-extra | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pairwise', round(d['value']), round(d['e2e']['value']), round(d['stage_ms_per_step']['dp'],2))"
done
