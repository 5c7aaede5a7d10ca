#!/usr/bin/env bash
# Round 2, call 17 (1 GPU): graph instantiated WITH node priorities (small kernels high, expansions low).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
for c in humanoid rocket robot_arm; do
  run 150 python tools/set_ab.py $c POCKIT_B200_GRAPH_PRIORITY=1,0
done
run 200 python tools/c5_probe.py POCKIT_B200_GRAPH_PRIORITY=1,0
