#!/bin/bash
# A/B of how the table build shares the SMs with the DP: stream priority (ISX_TABLES_PRIO) x resident table CTAs per
# SM (ISX_TAB_CTAS_PER_SM).  Prints resident frames/s per variant and workload.
tag=${1:-ab_tables}
out=gpurun_out/${tag}.txt
: > $out
for wl in unary_b64 pairwise_b64; do
  for v in "0 0" "1 1" "1 2" "0 1" "0 2"; do
    set -- $v
    r=$(ISX_TABLES_PRIO=$1 ISX_TAB_CTAS_PER_SM=$2 python bench.py --no-extra --no-cpu-baseline --workload $wl 2>/dev/null |
        python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print(round(d['value']), round(d['e2e']['value']), d['stage_ms_per_step'])")
    echo "$wl prio=$1 ctas=$2: $r" | tee -a $out
  done
done
