#!/bin/bash
# Same-box A/B: run the bench with each build/variants/libws_<name>.so, twice, interleaved.
out=gpurun_out; mkdir -p $out
for rep in 1 2; do
for name in "$@"; do
  WS_LIB_PATH=$PWD/build/variants/libws_$name.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-extra ${AB_ARGS} > $out/ab_$name.json 2> $out/ab_$name.err || tail -3 $out/ab_$name.err
  python - <<PY
import json
d=json.loads(open("$out/ab_$name.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("%-12s value %.1f e2e %.1f march %.3f merge %.3f replay %.3f reg %.3f step %.3f" % ("$name", d["value"], d["e2e"]["value"], k["march"], k["merge"], k["replay"], k["reg_20_iterations"], k["step_total"]))
PY
done; done
