#!/bin/bash
# Round-2 late experiments on one box: work-item size, grid shapes, cost of the profiling events, full captures of
# the merges / far-field march / replay.  Usage: bash tools/gpu_exp1.sh <tag>
tag=${1:-r02C}
out=gpurun_out; mkdir -p $out
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-extra > $out/${tag}_$name.json 2> $out/${tag}_$name.err || tail -3 $out/${tag}_$name.err
  python - <<PY
import json
d=json.loads(open("$out/${tag}_$name.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("%-14s value %.1f e2e %.1f upd %.3f march %.3f merge %.3f replay %.3f reg %.3f step %.3f" % ("$name", d["value"], d["e2e"]["value"], k["update_tsdf"], k["march"], k["merge"], k["replay"], k["reg_20_iterations"], k["step_total"]))
PY
}
B=$PWD/build/variants
for rep in 1 2; do
  run base$rep WS_LIB_PATH=$B/libws_base.so
  run blk32$rep WS_LIB_PATH=$B/libws_blk32.so
  run noprof$rep WS_LIB_PATH=$B/libws_base.so WS_BENCH_PROFILE=0
done
run g332 WS_LIB_PATH=$B/libws_base.so WS_LS_GRID_S=3 WS_LS_GRID_N=3 WS_LS_GRID_F=2
run g233 WS_LIB_PATH=$B/libws_base.so WS_LS_GRID_S=2 WS_LS_GRID_N=3 WS_LS_GRID_F=3
run g222 WS_LIB_PATH=$B/libws_base.so WS_LS_GRID_S=2 WS_LS_GRID_N=2 WS_LS_GRID_F=2
run g242 WS_LIB_PATH=$B/libws_base.so WS_LS_GRID_S=2 WS_LS_GRID_N=4 WS_LS_GRID_F=2
WS_LIB_PATH=$B/libws_base.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"merge_kernel|march_lockstep|replay" -s 21 -c 7 -f -o $out/${tag}_prof \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-extra > $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_ncu_full.log
ls -la $out | grep $tag | head -30
