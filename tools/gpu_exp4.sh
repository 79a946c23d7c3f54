#!/bin/bash
tag=${1:-r02H}
out=gpurun_out; mkdir -p $out
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-extra > $out/${tag}_$name.json 2> $out/${tag}_$name.err || tail -3 $out/${tag}_$name.err
  python - <<PY
import json
d=json.loads(open("$out/${tag}_$name.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("%-18s value %.1f e2e %.1f upd %.3f march %.3f merge %.3f replay %.3f reg %.3f step %.3f" % ("$name", d["value"], d["e2e"]["value"], k["update_tsdf"], k["march"], k["merge"], k["replay"], k["reg_20_iterations"], k["step_total"]))
PY
}
B=$PWD/build/variants
run c3_s2n3f3_k5 WS_LIB_PATH=$B/libws_split.so WS_NEAR_SPLIT=5 WS_LS_GRID_S=2 WS_LS_GRID_N=3 WS_LS_GRID_F=3
run f4_s2n4f4_k5 WS_LIB_PATH=$B/libws_free4.so WS_NEAR_SPLIT=5 WS_LS_GRID_S=2 WS_LS_GRID_N=4 WS_LS_GRID_F=4
run f4_s2n4f4_k8 WS_LIB_PATH=$B/libws_free4.so WS_NEAR_SPLIT=8 WS_LS_GRID_S=2 WS_LS_GRID_N=4 WS_LS_GRID_F=4
run f4_s2n3f4_k5 WS_LIB_PATH=$B/libws_free4.so WS_NEAR_SPLIT=5 WS_LS_GRID_S=2 WS_LS_GRID_N=3 WS_LS_GRID_F=4
run f4_s2n4f3_k5 WS_LIB_PATH=$B/libws_free4.so WS_NEAR_SPLIT=5 WS_LS_GRID_S=2 WS_LS_GRID_N=4 WS_LS_GRID_F=3
run a4_s2n4f4_k5 WS_LIB_PATH=$B/libws_all4.so WS_NEAR_SPLIT=5 WS_LS_GRID_S=2 WS_LS_GRID_N=4 WS_LS_GRID_F=4
run a4_s3n4f4_k5 WS_LIB_PATH=$B/libws_all4.so WS_NEAR_SPLIT=5 WS_LS_GRID_S=3 WS_LS_GRID_N=4 WS_LS_GRID_F=4
run a4_s4n4f4_k5 WS_LIB_PATH=$B/libws_all4.so WS_NEAR_SPLIT=5 WS_LS_GRID_S=4 WS_LS_GRID_N=4 WS_LS_GRID_F=4
run a4_s3n4f4_k8 WS_LIB_PATH=$B/libws_all4.so WS_NEAR_SPLIT=8 WS_LS_GRID_S=3 WS_LS_GRID_N=4 WS_LS_GRID_F=4
run a4_s3n3f4_k4 WS_LIB_PATH=$B/libws_all4.so WS_NEAR_SPLIT=4 WS_LS_GRID_S=3 WS_LS_GRID_N=3 WS_LS_GRID_F=4
run c3_s2n3f3_k5b WS_LIB_PATH=$B/libws_split.so WS_NEAR_SPLIT=5 WS_LS_GRID_S=2 WS_LS_GRID_N=3 WS_LS_GRID_F=3
