#!/usr/bin/env bash
# Round 2, call 24 (1 GPU): rolled staging loops in the column-walk / bulk kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 500 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py tests/test_gpu_zz_late.py -x -q -m gpu
for c in humanoid rocket robot_arm; do
  run 150 python tools/set_ab.py $c POCKIT_B200_SET=small
done
run 400 python bench.py --steps 20 --warmup 5 --no-all-configs --no-c5 --no-cpu-baseline --no-compact
