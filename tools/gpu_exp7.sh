#!/bin/bash
tag=${1:-r02P}; shift
out=gpurun_out; mkdir -p $out
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-extra > $out/${tag}_$name.json 2> $out/${tag}_$name.err || tail -3 $out/${tag}_$name.err
  python - <<PY
import json
d=json.loads(open("$out/${tag}_$name.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("%-12s value %.1f e2e %.1f upd %.3f march %.3f merge %.3f replay %.3f reg %.3f step %.3f" % ("$name", d["value"], d["e2e"]["value"], k["update_tsdf"], k["march"], k["merge"], k["replay"], k["reg_20_iterations"], k["step_total"]))
PY
}
B=$PWD/build/variants
timeout 1200 python -m pytest tests/test_full_size.py tests/test_gpu_parity.py tests/test_sharding.py tests/test_soak.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
for rep in 1 2; do
  for v in "$@"; do run $v$rep WS_LIB_PATH=$B/libws_$v.so; done
done
