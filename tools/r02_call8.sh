#!/usr/bin/env bash
# Round 2, call 8 (1 GPU): batch expansion kernel -- parity at B = 8192, stage times, set time.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 400 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -x -q -m gpu -k "quadrotor or batched or oracle"
run 200 python tools/stage_times.py quadrotor
run 200 python tools/c5_probe.py POCKIT_B200_EXPAND=,columns
