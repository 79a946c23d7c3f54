#!/bin/bash
tag=${1:-r02K}
out=gpurun_out; mkdir -p $out
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-extra > $out/${tag}_$name.json 2> $out/${tag}_$name.err || tail -3 $out/${tag}_$name.err
  python - <<PY
import json
d=json.loads(open("$out/${tag}_$name.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("%-18s value %.1f e2e %.1f upd %.3f march %.3f merge %.3f replay %.3f reg %.3f step %.3f" % ("$name", d["value"], d["e2e"]["value"], k["update_tsdf"], k["march"], k["merge"], k["replay"], k["reg_20_iterations"], k["step_total"]))
for r in d["roofline"]["update_timeline_ms"][5:8]: print("     %-62s %.4f %.4f (%.3f)" % (r["range"], r["start"], r["stop"], r["stop"]-r["start"]))
PY
}
B=$PWD/build/variants
timeout 900 python -m pytest tests/test_full_size.py tests/test_gpu_parity.py tests/test_sharding.py tests/test_soak.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
run f2_1 WS_LIB_PATH=$B/libws_fab.so WS_LS_GRID_F=2 WS_LS_GRID_F2=1
run f3_0 WS_LIB_PATH=$B/libws_fab.so WS_LS_GRID_F=3 WS_LS_GRID_F2=0
run f2_0 WS_LIB_PATH=$B/libws_fab.so WS_LS_GRID_F=2 WS_LS_GRID_F2=0
run f2_2 WS_LIB_PATH=$B/libws_fab.so WS_LS_GRID_F=2 WS_LS_GRID_F2=2
run f2_1_k6 WS_LIB_PATH=$B/libws_fab.so WS_LS_GRID_F=2 WS_LS_GRID_F2=1 WS_NEAR_SPLIT=6
run f2_1_k4 WS_LIB_PATH=$B/libws_fab.so WS_LS_GRID_F=2 WS_LS_GRID_F2=1 WS_NEAR_SPLIT=4
run f2_1b WS_LIB_PATH=$B/libws_fab.so WS_LS_GRID_F=2 WS_LS_GRID_F2=1
