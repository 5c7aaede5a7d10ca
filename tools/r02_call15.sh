#!/usr/bin/env bash
# Round 2, call 15 (1 GPU): bulk-store expansion again on the LGL configurations (after streaming stores).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
for c in humanoid rocket robot_arm; do
  run 150 python tools/set_ab.py $c POCKIT_B200_EXPAND=params,bulk
done
