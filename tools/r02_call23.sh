#!/usr/bin/env bash
# Round 2, call 23 (1 GPU): defect jobs of all phases in one launch; weighted-shard GPU test; whole tier.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 700 python -m pytest tests -m gpu -x -q
for c in rocket humanoid robot_arm; do
  run 150 python tools/set_ab.py $c POCKIT_B200_SET=small
done
