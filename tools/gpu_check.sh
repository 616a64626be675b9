(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3); timeout 300 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_unary.log 2>&1; timeout 300 python bench.py --workload pairwise_b64 --no-cpu-baseline --no-extra > gpurun_out/bench_pairwise.log 2>&1; python - <<EOF
import json
for f in ("unary","pairwise"):
    try:
        d=json.loads(open("gpurun_out/bench_%s.log"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k:round(v,2) for k,v in d["stage_ms_per_step"].items()}, round(d["roofline"]["frac"],3))
    except Exception as e: print(f, "ERR", e, open("gpurun_out/bench_%s.log"%f).read()[-800:])
EOF
