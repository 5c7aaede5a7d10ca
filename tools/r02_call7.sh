#!/usr/bin/env bash
# Round 2, call 7 (1 GPU): defect kernel without staging (tests + stage times), ncu --set full of one whole set.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -x -q -m gpu
run 200 python tools/stage_times.py quadrotor humanoid robot_arm
skip() { case $1 in robot_arm) echo 24;; *) echo 28;; esac; }
for c in robot_arm humanoid quadrotor; do
  echo "### ncu full $c"
  timeout 600 ncu --set full --clock-control none --import-source on -s $(skip $c) -c 14 -f -o gpurun_out/r02_full_$c \
    python tools/ncu_target.py $c 3 > gpurun_out/r02_ncu_$c.log 2>&1
  echo "### exit $?"
done
ls -la gpurun_out/*.ncu-rep
