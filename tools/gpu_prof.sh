#!/bin/bash
# ncu --set full capture of kernels matching a regex.  Usage: bash tools/gpu_prof.sh <tag> <regex> [count] [extra bench args]
tag=$1; rx=$2; cnt=${3:-2}; shift 3
out=gpurun_out
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 6 -c $cnt -f -o $out/${tag}_prof \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda "$@" > $out/${tag}_ncu_full.log 2>&1
tail -3 $out/${tag}_ncu_full.log
ls -la $out/${tag}_prof.ncu-rep
