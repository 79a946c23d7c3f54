#!/bin/bash
# Quick GPU visit: parity tests + bench (no CPU baseline / reference CUDA legs).  Usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
tag=${1:-q}
out=gpurun_out
mkdir -p $out
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > $out/${tag}_pytest.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
fi
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_scan"], d["roofline"]["frac"])
print(d["work"])
PY
