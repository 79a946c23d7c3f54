#!/bin/bash
# registration A/B: builds x WS_REG_BLOCKS
out=gpurun_out; mkdir -p $out
run() { name=$1; blocks=$2
  WS_REG_BLOCKS=$blocks WS_LIB_PATH=$PWD/build/variants/libws_$name.so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-extra > $out/abr_${name}_$blocks.json 2> $out/abr_${name}_$blocks.err || tail -3 $out/abr_${name}_$blocks.err
  python - <<PY
import json
d=json.loads(open("$out/abr_${name}_$blocks.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernel_ms_per_scan"]
print("%-8s blocks %4s value %.1f reg %.4f step %.3f" % ("$name", "$blocks", d["value"], k["reg_20_iterations"], k["step_total"]))
PY
}
run base 999; run base 148; run base 222; run base 74; run p4 148; run p4 74; run t512 999; run t512 148; run t512 74; run t1024 148; run t1024 74
