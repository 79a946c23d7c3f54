#!/bin/bash
# bench only (guarded).  Usage: bash tools/gpu_b.sh <tag> [extra bench args]
tag=${1:-r02x}; shift
out=gpurun_out; mkdir -p $out
timeout 280 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda "$@" > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("value %.1f e2e %.1f update %.3f (march %.3f merge %.3f replay %.3f) reg %.3f step %.3f frac %.4f" % (d["value"], d["e2e"]["value"], k["update_tsdf"], k["march"], k["merge"], k["replay"], k["reg_20_iterations"], k["step_total"], d["roofline"]["frac"]))
print(d["work"]); print(d.get("sub_configs")); print(d.get("parity_check"))
PY
tail -3 $out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
for r in d["roofline"].get("update_timeline_ms", []): print("  %-50s %8.3f -> %8.3f" % (r["range"], r["start"], r["stop"]))
PY
