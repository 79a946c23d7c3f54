#!/bin/bash
# bench at N GPUs for several stripe widths (0 = contiguous slabs).  Usage: bash tools/gpu_stripes.sh <tag> <N> <widths...>
tag=$1; N=$2; shift 2
out=gpurun_out; mkdir -p $out
for w in "$@"; do
  WS_STRIPE_COLS=$w timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2970$N bench.py --gpus $N --steps 30 --warmup 5 > $out/${tag}_n${N}_w$w.json 2> $out/${tag}_n${N}_w$w.err || tail -3 $out/${tag}_n${N}_w$w.err
  python - <<PY
import json
d=json.loads(open("$out/${tag}_n${N}_w$w.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("N=$N stripe=$w value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {a: round(b,3) for a,b in k.items()}, "T", d["work"]["T"])
PY
done
