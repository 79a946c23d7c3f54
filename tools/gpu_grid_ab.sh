#!/bin/bash
# A/B of launch geometry knobs: lines of "ENV=VAL ..." are tried one after the other
while read -r cfg; do
  [ -z "$cfg" ] && continue
  echo "== $cfg"
  env $cfg bash tools/gpu_b.sh r02ab --no-extra 2>&1 | grep "^value\|exit\|replay rounds\|free-space merge"
done <<CFGS
WS_REPLAY_CTAS=2 WS_FMERGE_GRID=4
WS_REPLAY_CTAS=1 WS_FMERGE_GRID=6
WS_REPLAY_CTAS=1 WS_FMERGE_GRID=4
WS_REPLAY_CTAS=2 WS_FMERGE_GRID=3
WS_REPLAY_CTAS=4 WS_FMERGE_GRID=8
CFGS
