#!/bin/bash
for cfg in "1 1" "2 1" "1 2" "2 2"; do
  set -- $cfg
  echo "== S=$1 N=$2"
  WS_LS_GRID_S=$1 WS_LS_GRID_N=$2 bash tools/gpu_b.sh r02B_g$1$2 --no-extra 2>&1 | grep "^value\|exit"
done
