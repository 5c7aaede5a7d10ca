#!/usr/bin/env bash
# Round 2, call 25 (1 GPU): slot-order batch expansion (pk_expand_slots) -- parity, then A/B on configs[4].
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 300 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "quadrotor"
run 200 python tools/c5_probe.py POCKIT_B200_EXPAND=,slots
POCKIT_B200_EXPAND=slots run 200 python tools/c5_probe.py POCKIT_B200_SLOT_LISTS=1,3,4,8
POCKIT_B200_EXPAND=slots run 100 python tools/stage_times.py quadrotor
