#!/usr/bin/env bash
# Round 2, call 3 (1 GPU): staggered set schedule A/B, re-run of the solver tests.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 300 python -m pytest tests/test_gpu_solver.py tests/test_gpu_outputs.py tests/test_gpu_parity.py -x -q -m gpu
for c in robot_arm humanoid rocket; do
  run 150 python tools/set_ab.py $c POCKIT_B200_STAGGER=0,1,2
done
run 150 env CUDA_DEVICE_MAX_CONNECTIONS=32 python tools/set_ab.py humanoid POCKIT_B200_STAGGER=0,2
run 150 env POCKIT_B200_EXPAND=bulk python tools/set_ab.py humanoid POCKIT_B200_STAGGER=2
run 150 env POCKIT_B200_EXPAND=bulk python tools/set_ab.py rocket POCKIT_B200_STAGGER=2
