#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of the main kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py --steps 40 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
tail -1 $out/${tag}_bench.json
timeout 300 python bench.py --steps 40 --warmup 5 --update-only --no-cpu-baseline --no-ref-cuda > $out/${tag}_bench_update_only.json 2>> $out/${tag}_bench.err
tail -1 $out/${tag}_bench_update_only.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file $out/${tag}_launches.csv \
   python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-cuda > $out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"setup_kernel|march|merge|brick_list|replay|reg_loop" -s 36 -c 12 -f -o $out/${tag}_prof \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ref-cuda > $out/${tag}_ncu_full.log 2>&1
ls -la $out
