#!/bin/bash
# Scaling visit (gpurun --gpus 8): bench at N=1,2,4,8 on the default workload, then one short config-3 run.
tag=$1
out=gpurun_out; mkdir -p $out
nvidia-smi -L | wc -l
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-ref-cuda > $out/${tag}_n1.json 2> $out/${tag}_n1.err
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n bench.py --gpus $n --steps 30 --warmup 5 > $out/${tag}_n$n.json 2> $out/${tag}_n$n.err; echo "bench n=$n exit $?"
  grep -v "OMP_NUM_THREADS\|^\*\*\*" $out/${tag}_n$n.err | tail -2
done
# config 3: 128x2048 scan, 2048^3 @ 2 cm over 8 GPUs (short)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 8 --steps 6 --warmup 3 --grid 2048 --res 20 --cols 2048 > $out/${tag}_cfg3.json 2> $out/${tag}_cfg3.err; echo "config3 exit $?"
grep -v "OMP_NUM_THREADS\|^\*\*\*" $out/${tag}_cfg3.err | tail -4
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernel_ms_per_scan"]
        print(f.split("/")[-1], "gpus", d["n_gpus"], "value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {a: round(b,3) for a,b in k.items()}, "T", d["work"]["T"], "C", d["work"]["C"])
    except Exception as e:
        print(f, "ERR", e)
PY
