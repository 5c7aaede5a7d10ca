#!/usr/bin/env bash
# Round 2, call 26 (1 GPU): batch expansion instantiated per row count (5 unrolled rows instead of 16),
# lists per thread as a launch argument -- parity, then A/B on configs[4].
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 300 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -x -q -m gpu -k "quadrotor or batched"
run 200 python tools/c5_probe.py POCKIT_B200_BATCH_LISTS=,1,3,5
run 100 python tools/stage_times.py quadrotor
