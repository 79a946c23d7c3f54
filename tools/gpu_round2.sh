#!/bin/bash
# One GPU-box visit: parity tests, the default bench line, ncu launch list + per-kernel metrics + full capture.
# Usage (under gpurun): bash tools/gpu_round2.sh <tag>
tag=${1:-r02z}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
tail -c 600 $out/${tag}_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 200 --csv --log-file $out/${tag}_launches.csv \
   python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-extra > $out/${tag}_ncu_launch.log 2>&1
bash tools/gpu_list.sh $tag --no-extra > $out/${tag}_list.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"setup_kernel|item_scan|march|merge|replay|reg_loop" -s 22 -c 11 -f -o $out/${tag}_prof \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-extra > $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_ncu_full.log
ls -la $out | grep $tag
timeout 600 python tests/soak.py 60 4242 > $out/${tag}_soak.log 2>&1; tail -2 $out/${tag}_soak.log
