#!/bin/bash
# Quick GPU visit: parity tests, short bench, one ncu --set full capture of the march.  Usage: bash tools/gpu_quick2.sh <tag> [kernel regex]
tag=${1:-r02x}; rx=${2:-march_lockstep}
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-ref-cuda > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
tail -c 3000 $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 4 -c 1 -f -o $out/${tag}_prof \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > $out/${tag}_ncu_full.log 2>&1
tail -3 $out/${tag}_ncu_full.log
