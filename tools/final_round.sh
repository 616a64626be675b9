#!/bin/bash
# Final evidence of the round on one GPU: the whole GPU test suite, the default bench line (all workloads), the
# reference arm for the three batch-64 configurations (+ the ncu launch list of its unary run: which kernels and copies
# the reference's 9.6 ms per frame are made of), smoke().
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_final_pytest.log
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
: > gpurun_out/r2_final_ref.jsonl
for w in unary_b64 pairwise_b64 pairwise_w4_b64; do
  python bench.py --impl reference --workload $w --steps 5 --warmup 1 >> gpurun_out/r2_final_ref.jsonl 2>> gpurun_out/r2_final_bench.err
done
for w in unary_b64 pairwise_b64; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_ref_launches_$w.csv \
    python bench.py --impl reference --workload $w --steps 1 --warmup 1 > gpurun_out/r2_ref_launches_$w.log 2>&1
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
cat gpurun_out/r2_final_pytest.log gpurun_out/r2_final_smoke.log
