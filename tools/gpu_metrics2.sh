#!/bin/bash
# ncu metrics ($METRICS) of kernels matching $1 for each variant in the remaining args, full bench (with registration)
rx=$1; shift
out=gpurun_out; mkdir -p $out
M=${METRICS:-gpu__time_duration.sum,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_active.avg,smsp__inst_executed.min,smsp__inst_executed.max,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max,launch__grid_size,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,smsp__cycles_active.min,smsp__cycles_active.max}
for name in "$@"; do
  WS_LIB_PATH=$PWD/build/variants/libws_$name.so timeout 600 ncu --metrics $M --clock-control none -k regex:"$rx" -s 4 -c 1 --csv --log-file $out/met_$name.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > /dev/null 2>&1
  echo "== $name"; grep -v "^==" $out/met_$name.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
for r in rows[1:]:
    print(r[h.index('Metric Name')], r[h.index('Metric Value')])
"
done
