#!/bin/bash
# The five BASELINE.json configurations on one box: one bench.py JSON line per configuration and arm.
# Usage: bash tools/run_configs.sh <tag> [N gpus]   -> gpurun_out/<tag>_configs.jsonl
tag=${1:-r1}; N=${2:-1}
out=gpurun_out/${tag}_configs_n${N}.jsonl; mkdir -p gpurun_out; : > $out
run() {
  if [ "$N" = 1 ]; then timeout 900 python bench.py --gpus 1 "$@"
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; fi
}
# configs[0] + [1]: unary; the line carries the CPU port (cpu_baseline) and the batch-1 latency
run --workload unary_b64 2>gpurun_out/${tag}_err.log | tail -1 >> $out
run --impl reference --workload unary_b64 --steps 3 2>>gpurun_out/${tag}_err.log | tail -1 >> $out
# configs[2]: pairwise + instance grouping
run --workload pairwise_b64 --no-cpu-baseline 2>>gpurun_out/${tag}_err.log | tail -1 >> $out
run --impl reference --workload pairwise_b64 --steps 3 2>>gpurun_out/${tag}_err.log | tail -1 >> $out
# configs[3]: stixel width 4
run --workload pairwise_w4_b64 --no-cpu-baseline 2>>gpurun_out/${tag}_err.log | tail -1 >> $out
run --impl reference --workload pairwise_w4_b64 --steps 3 2>>gpurun_out/${tag}_err.log | tail -1 >> $out
# configs[4]: the 4096-frame stream at batch 256 (4096 / (256 N) steps per rank)
run --workload pairwise_stream_b256 --steps $((16 / N)) --no-cpu-baseline 2>>gpurun_out/${tag}_err.log | tail -1 >> $out
python - <<PY
import json
for l in open("$out"):
    try:
        d = json.loads(l)
        print(d["config"]["workload"], d.get("impl", "ours"), "n=%d" % d["n_gpus"], "value %.0f" % d["value"],
              "e2e %.0f" % d["e2e"]["value"], "frac %.3f" % d.get("roofline", {}).get("frac", 0),
              d.get("latency_ms_batch1"), d.get("clocks"))
    except Exception as e:
        print("bad line", e, l[:200])
PY
tail -5 gpurun_out/${tag}_err.log
