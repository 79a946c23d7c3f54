#!/bin/bash
tag=${1:-r02E}
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
  run base_p1_$rep WS_LIB_PATH=$B/libws_base.so WS_BENCH_PROFILE=1
  run base_p0_$rep WS_LIB_PATH=$B/libws_base.so WS_BENCH_PROFILE=0
  run new_p2_$rep WS_LIB_PATH=$B/libws_new.so WS_BENCH_PROFILE=2
  run new_p0_$rep WS_LIB_PATH=$B/libws_new.so WS_BENCH_PROFILE=0
  run newinl_p0_$rep WS_LIB_PATH=$B/libws_new.so WS_BENCH_PROFILE=0 WS_GENERAL_INLINE=1
done
