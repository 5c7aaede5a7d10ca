#!/usr/bin/env bash
# Round 2, call 16 (1 GPU): bulk-store expansion as the default for unaligned blocks, phases' node programs side by side.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 700 python -m pytest tests -m gpu -x -q
for c in rocket humanoid robot_arm; do
  run 150 python tools/set_ab.py $c POCKIT_B200_SET=small
done
