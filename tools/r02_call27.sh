#!/usr/bin/env bash
# Round 2, call 27 (1 GPU): does the batch expansion want FEWER resident blocks?  (call 26: 32 registers
# instead of 48 made the Jacobian expansion slower, 68 -> 83 us.)  Unused dynamic shared memory caps the
# blocks per SM: 0 -> 16, 18000 -> 12, 28000 -> 8, 37000 -> 6, 56000 -> 4, 75000 -> 3.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
echo "== 16 unrolled rows (48 registers, 10 blocks per SM uncapped)"
run 200 python tools/c5_probe.py POCKIT_B200_BATCH_SMEM=,28000,37000,56000,75000
echo "== exact rows (32 registers, 16 blocks per SM uncapped)"
POCKIT_B200_BATCH_ROWS=exact run 200 python tools/c5_probe.py POCKIT_B200_BATCH_SMEM=18000,28000,37000,56000,75000
echo "== slot order, 5 lists per thread"
POCKIT_B200_EXPAND=slots POCKIT_B200_SLOT_LISTS=5 run 200 python tools/c5_probe.py POCKIT_B200_BATCH_SMEM=18000,28000,47000
