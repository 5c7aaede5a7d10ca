#!/usr/bin/env bash
# Round 2, call 28 (1 GPU): ncu --set full of the batch expansion launches only (quadrotor B = 8192), three
# shapes: default (16 unrolled rows), exact rows, slot order.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() { # name, env...
  local name=$1; shift
  echo "### ncu full, batch expansion only: $name"
  env "$@" timeout 300 ncu --set full --clock-control none --import-source on -k regex:pk_expand -c 6 -f -o gpurun_out/r02_full_batch_$name \
    python tools/ncu_target.py quadrotor 3 > gpurun_out/r02_ncu_batch_$name.log 2>&1
  echo "### exit $?"
}
cap default POCKIT_B200_NOP=1
cap exact_rows POCKIT_B200_BATCH_ROWS=exact
cap slots POCKIT_B200_EXPAND=slots POCKIT_B200_SLOT_LISTS=5
