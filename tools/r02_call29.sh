#!/usr/bin/env bash
# Round 2, call 29 (1 GPU): batch expansion kernels with the list values loaded ahead of the stores, the list
# groups from a parameter table, streaming stores decided at compile time -- parity, then A/B on configs[4].
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 300 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -x -q -m gpu -k "quadrotor or batched"
echo "== column mapping, 16 unrolled rows"
run 200 python tools/c5_probe.py POCKIT_B200_BATCH_LISTS=,4,8
echo "== column mapping, exact rows"
POCKIT_B200_BATCH_ROWS=exact run 200 python tools/c5_probe.py POCKIT_B200_BATCH_LISTS=2,4,8
echo "== slot order"
POCKIT_B200_EXPAND=slots run 200 python tools/c5_probe.py POCKIT_B200_SLOT_LISTS=2,4,8
