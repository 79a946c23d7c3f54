#!/bin/bash
# A/B of launch geometry knobs / library variants: lines of "ENV=VAL ..." are tried one after the other
while read -r cfg; do
  [ -z "$cfg" ] && continue
  echo "== $cfg"
  env $cfg bash tools/gpu_b.sh r02ab --no-extra 2>&1 | grep "^value\|exit [1-9]\|surface march\|near field\|far field\|record pass"
done <<CFGS
WS_X=0
WS_LS_GRID_F=3
WS_LS_GRID_N=3
WS_LS_GRID_S=3 WS_LS_GRID_N=3 WS_LS_GRID_F=3
WS_LS_GRID_F=1
CFGS
