#!/usr/bin/env bash
# Round 2, call 13 (1 GPU): batch expansion kernel v2 (P in registers, all lists of a job per thread).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 400 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -x -q -m gpu -k "quadrotor or batched"
run 200 python tools/stage_times.py quadrotor
run 200 python tools/c5_probe.py POCKIT_B200_EXPAND=,columns
