#!/bin/bash
# a few ncu metrics of kernels matching $1 for each variant in the remaining args
rx=$1; shift
out=gpurun_out; mkdir -p $out
for name in "$@"; do
  WS_LIB_PATH=$PWD/build/variants/libws_$name.so timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"$rx" -s 4 -c 2 --csv --log-file $out/met_$name.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --update-only > /dev/null 2>&1
  echo "== $name"; grep -v "^==" $out/met_$name.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
for r in rows[1:]:
    print(r[h.index('Kernel Name')][:40], r[h.index('Metric Name')], r[h.index('Metric Value')])
"
done
