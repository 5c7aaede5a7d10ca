#!/usr/bin/env bash
# Round 2, call 19 (1 GPU): final ncu --set full captures of one whole set, humanoid (TMA bulk-store expansion) and quadrotor batch.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in humanoid quadrotor; do
  echo "### ncu full $c"
  timeout 600 ncu --set full --clock-control none --import-source on -s 24 -c 12 -f -o gpurun_out/r02_full_$c \
    python tools/ncu_target.py $c 3 > gpurun_out/r02_ncu_$c.log 2>&1
  echo "### exit $?"
done
