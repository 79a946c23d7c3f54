#!/bin/bash
# ncu --set full of kernels matching a regex (launch-skip 2 matching, count N).  Usage: bash tools/gpu_prof2.sh <tag> <regex> [count]
tag=$1; rx=$2; cnt=${3:-1}
out=gpurun_out; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 3 -c $cnt -f -o $out/${tag}_prof \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_ncu_full.log
