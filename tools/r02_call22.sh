#!/usr/bin/env bash
# Round 2, call 22 (1 GPU): pk_eval_set starts the Jacobian before the multipliers are uploaded; whole tier + bench.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 700 python -m pytest tests -m gpu -x -q
run 700 python bench.py --steps 20 --warmup 5 --no-all-configs --no-c5
