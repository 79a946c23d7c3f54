#!/bin/bash
# guarded GPU run: smoke first under a short timeout (a hang costs minutes, not the round), then tests + bench + list.
# Usage: bash tools/gpu_safe.sh <tag>
tag=${1:-r02x}
out=gpurun_out; mkdir -p $out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; rc=$?
echo "smoke exit $rc"; tail -3 $out/${tag}_smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 200 python -m pytest tests/test_cpp_shim.py -m gpu -x -q --timeout 150 > $out/${tag}_shim.log 2>&1; rc=$?
echo "shim exit $rc"; tail -3 $out/${tag}_shim.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1000 python -m pytest tests -m gpu -x -q --timeout 240 > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -8 $out/${tag}_pytest.log
timeout 240 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("value %.1f e2e %.1f update %.3f (march %.3f merge %.3f replay %.3f) reg %.3f step %.3f frac %.4f" % (d["value"], d["e2e"]["value"], k["update_tsdf"], k["march"], k["merge"], k["replay"], k["reg_20_iterations"], k["step_total"], d["roofline"]["frac"]))
print(d["work"])
PY
tail -3 $out/${tag}_bench.err
timeout 300 bash tools/gpu_list.sh $tag 2>&1 | head -16
