#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): two-process peer check, then bench at 1..N GPUs.  Usage: bash tools/gpu_multi.sh <tag> <N>
tag=$1; N=${2:-2}
out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mp_peer_check.py > $out/${tag}_mpcheck.log 2>&1; echo "mp check exit $?"; tail -3 $out/${tag}_mpcheck.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
n=2
while [ $n -le $N ]; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 40 --warmup 5 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err; echo "bench n=$n exit $?"
  tail -2 $out/${tag}_bench_n$n.err
  n=$((n*2))
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$out/${tag}_bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernel_ms_per_scan"]
        print(f, "gpus", d["n_gpus"], "value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {a: round(b,3) for a,b in k.items()}, d["work"]["T"], d["work"]["C"])
    except Exception as e:
        print(f, "ERR", e)
PY
