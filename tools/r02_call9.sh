#!/usr/bin/env bash
# Round 2, call 9 (1 GPU): the five-callback pipeline again (one per-node program for everything), with launch lists.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
for c in robot_arm humanoid rocket; do
  run 150 python tools/set_ab.py $c POCKIT_B200_SET=small,1
done
for c in humanoid robot_arm; do
  echo "### launch list SET=1 $c"
  POCKIT_B200_SET=1 timeout 300 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none -c 200 --csv \
    --log-file gpurun_out/r02_launches_set1_$c.csv python tools/ncu_target.py $c 3 > /dev/null 2>&1
  echo "### exit $?"
done
