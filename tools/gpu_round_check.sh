#!/bin/bash
# One GPU call: parity tests, the two main bench lines, the ncu launch list of the bench command and a full-set
# capture of one 16-frame chunk per mode.  Usage: bash tools/gpu_round_check.sh <tag>
tag=${1:-r1g}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_unary.json 2> gpurun_out/${tag}_bench_unary.err
timeout 600 python bench.py --workload pairwise_b64 --no-cpu-baseline > gpurun_out/${tag}_bench_pairwise.json 2> gpurun_out/${tag}_bench_pairwise.err
timeout 300 python bench.py --impl reference --steps 3 > gpurun_out/${tag}_bench_ref_unary.json 2> gpurun_out/${tag}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_unary_b64.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_launches_bench.log 2>&1
for mode in unary pairwise; do
  timeout 900 ncu --set full --clock-control none --import-source on -c 12 -f -o gpurun_out/prof_${tag}_${mode} \
    python tools/profile_run.py --mode $mode --batch 32 --steps 1 > gpurun_out/prof_${tag}_${mode}.log 2>&1
done
cat gpurun_out/${tag}_pytest.log; tail -c 1500 gpurun_out/${tag}_bench_unary.json; tail -c 600 gpurun_out/${tag}_bench_unary.err
ls -la gpurun_out
