#!/usr/bin/env bash
# Round 2, call 18 (1 GPU): SCALED fast path; whole tier; set times.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 700 python -m pytest tests -m gpu -x -q
run 200 python tools/stage_times.py humanoid
for c in humanoid rocket robot_arm; do
  run 150 python tools/set_ab.py $c POCKIT_B200_SET=small
done
