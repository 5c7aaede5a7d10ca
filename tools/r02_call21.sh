#!/usr/bin/env bash
# Round 2, call 21 (1 GPU): reciprocal division in the column-walk / bulk kernels; five-callback pipeline on the batch.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 400 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -x -q -m gpu
for c in humanoid rocket robot_arm; do
  run 150 python tools/set_ab.py $c POCKIT_B200_SET=small
done
run 300 python tools/c5_probe.py POCKIT_B200_SET=small,1
