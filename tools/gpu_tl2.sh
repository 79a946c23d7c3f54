#!/bin/bash
# parity tests (default build + a variant that forces several flag-compaction rounds per merge CTA) + short bench + per-kernel list.
# Usage: bash tools/gpu_tl2.sh <tag>
tag=${1:-r02x}
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -8 $out/${tag}_pytest.log
if [ -f build/variants/libws_cap256.so ]; then
  WS_LIB_PATH=$PWD/build/variants/libws_cap256.so timeout 900 python -m pytest tests/test_full_size.py tests/test_gpu_parity.py tests/test_sharding.py -m gpu -x -q > $out/${tag}_pytest_cap256.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_cap256.log
  tail -3 $out/${tag}_pytest_cap256.log
fi
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("value %.1f e2e %.1f upd %.3f march %.3f merge %.3f replay %.3f reg %.3f step %.3f" % (d["value"], d["e2e"]["value"], k["update_tsdf"], k["march"], k["merge"], k["replay"], k["reg_20_iterations"], k["step_total"]))
print(d["work"])
for r in d["roofline"]["update_timeline_ms"]: print("  %-62s %.4f %.4f (%.3f)" % (r["range"], r["start"], r["stop"], r["stop"]-r["start"]))
print({k: (v["value"], v["e2e"]) for k, v in d.get("extra", {}).get("configs", {}).items()})
PY
tail -3 $out/${tag}_bench.err
bash tools/gpu_list.sh $tag 2>&1 | head -16
