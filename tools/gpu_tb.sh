#!/bin/bash
# parity tests + full default bench line.  Usage: bash tools/gpu_tb.sh <tag> [bench args]
tag=${1:-r02x}; shift
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -12 $out/${tag}_pytest.log
timeout 900 python bench.py "$@" > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
    k=d["roofline"]["kernel_ms_per_scan"]
    print("value %.1f e2e %.1f march %.3f merge %.3f replay %.3f reg %.3f step %.3f frac %.4f" % (d["value"], d["e2e"]["value"], k["march"], k["merge"], k["replay"], k["reg_20_iterations"], k["step_total"], d["roofline"]["frac"]))
    print(d["work"])
    for kk,v in d.get("extra",{}).get("configs",{}).items(): print(kk, "value %.1f e2e %.1f"%(v["value"],v["e2e"]), v["kernel_ms_per_scan"], v["work"])
    print("cpu", d.get("cpu_baseline",{}).get("value"), "refcuda", d.get("reference_cuda_same_gpu",{}).get("value"))
except Exception as e: print("parse failed", e)
PY
tail -5 $out/${tag}_bench.err
