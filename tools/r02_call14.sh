#!/usr/bin/env bash
# Round 2, call 14 (2 GPUs): staged asynchronous evaluation (tests), two-phase mesh-shard protocol, bench N = 2.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 400 python -m pytest tests/test_gpu_outputs.py tests/test_gpu_zz_late.py -x -q -m gpu
bash tools/r02_multi.sh 2 20
