#!/usr/bin/env bash
# Round 2, call 12 (1 GPU): whole GPU tier after the small-kernel work, stage times, set times.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 600 python -m pytest tests -m gpu -x -q
run 200 python tools/stage_times.py humanoid rocket
for c in robot_arm humanoid rocket; do
  run 150 python tools/set_ab.py $c POCKIT_B200_SET=small
done
run 200 python tools/c5_probe.py POCKIT_B200_BATCH_TABLES=0
