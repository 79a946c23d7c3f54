#!/bin/bash
# what the driver does at round end: smoke(), pytest -m gpu, the default bench line, the reference arm (short)
out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > $out/final_bench.json 2> $out/final_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("$out/final_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","scaling","dtype","gpu_launches")})
print("e2e", d["e2e"]["value"], "clocks", d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved","peak","frac","traffic")}, d["roofline"]["kernel_ms_per_scan"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "refcuda", d["reference_cuda_same_gpu"].get("value"))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-200
