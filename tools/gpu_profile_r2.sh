#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) full-set capture of every kernel of one
# launch-sized chunk (unary 32 frames, pairwise 64).  Usage: bash tools/gpu_profile_r2.sh <tag>
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_unary_b64.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_launches_bench.log 2>&1
# one launch of each mode's size: 32 frames in unary mode, 64 in pairwise mode
for mode in unary pairwise; do
  batch=32; [ $mode = pairwise ] && batch=64
  timeout 900 ncu --set full --clock-control none --import-source on -c 12 -f -o gpurun_out/prof_${tag}_${mode} \
    python tools/profile_run.py --mode $mode --batch $batch --steps 1 > gpurun_out/prof_${tag}_${mode}.log 2>&1
done
