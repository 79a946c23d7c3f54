#!/bin/bash
# 2-GPU visit: two-process peer check, the 2-GPU parity test, bench at N=2 (driver launch line).  Usage: bash tools/gpu_n2.sh <tag>
tag=$1; out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_sharding.py -m gpu -x -q > $out/${tag}_pytest_shard.log 2>&1; tail -2 $out/${tag}_pytest_shard.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 3 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err; echo "bench n=2 exit $?"
tail -3 $out/${tag}_bench_n2.err
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_n2.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("gpus", d["n_gpus"], "value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {a: (round(b,3) if isinstance(b,float) else b) for a,b in k.items() if a!="note"})
print(d.get("parity_check")); print(d["roofline"].get("kernel_ms_per_scan_min_over_ranks"))
PY
