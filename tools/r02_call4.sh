#!/usr/bin/env bash
# Round 2, call 4 (1 GPU): streaming stores -- whole GPU tier, A/B, stage times of the batched config, bench line.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 600 python -m pytest tests -m gpu -x -q
for c in robot_arm humanoid rocket; do
  run 150 python tools/set_ab.py $c POCKIT_B200_STREAM=1,0
done
run 150 python tools/stage_times.py quadrotor
run 500 python bench.py --steps 20 --warmup 5
