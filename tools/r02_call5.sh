#!/usr/bin/env bash
# Round 2, call 5 (1 GPU): small set pipeline (objective+gradient+constraints as one chain) A/B + tests.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_outputs.py -x -q -m gpu
for c in robot_arm humanoid rocket; do
  run 150 python tools/set_ab.py $c POCKIT_B200_SET=0,small
done
