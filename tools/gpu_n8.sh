#!/bin/bash
# 8-GPU visit with the driver's launch line: default workload at N=8 (parity check against one GPU + configs[3] in `extra`), then N=4.
tag=$1; out=gpurun_out; mkdir -p $out
nvidia-smi -L | wc -l
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2965$n bench.py --gpus $n --steps 20 --warmup 3 > $out/${tag}_n$n.json 2> $out/${tag}_n$n.err; echo "bench n=$n exit $?"
  grep -v "OMP_NUM_THREADS\|^\*\*\*" $out/${tag}_n$n.err | tail -3
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernel_ms_per_scan"]
        print(f.split("/")[-1], "gpus", d["n_gpus"], "value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {a: (round(b,3) if isinstance(b,float) else b) for a,b in k.items() if a!="note"})
        print("   min over ranks", d["roofline"].get("kernel_ms_per_scan_min_over_ranks"))
        print("   parity", (d.get("parity_check") or {}).get("vs_single_gpu"), (d.get("parity_check") or {}).get("scans"))
        for kk, v in (d.get("extra", {}).get("configs", {}) or {}).items():
            print("   ", kk, "value %.1f e2e %.1f" % (v["value"], v["e2e"]), v["kernel_ms_per_scan"], v["work"])
    except Exception as e:
        print(f, "ERR", e)
PY
