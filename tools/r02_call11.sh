#!/usr/bin/env bash
# Round 2, call 11 (1 GPU): reciprocal multipliers in defect / persistent-expansion kernels; bench line with the sustained roofline.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 600 python -m pytest tests -m gpu -x -q
run 200 python tools/stage_times.py quadrotor
run 200 python tools/c5_probe.py POCKIT_B200_BATCH_TABLES=0
run 150 python tools/set_ab.py humanoid POCKIT_B200_SET=small
run 600 python bench.py --steps 20 --warmup 5
