#!/usr/bin/env bash
# Round 2, call 6 (1 GPU): ncu --set full of every kernel of one evaluation set, three configurations.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in robot_arm humanoid quadrotor; do
  echo "### ncu $c"
  # launch list first (names + durations), then the full set of the second evaluation set
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_$c.csv \
    python tools/ncu_target.py $c 3 > gpurun_out/r02_ncu_$c.log 2>&1
  echo "### exit $?"
  timeout 500 ncu --set full --clock-control none --import-source on -s 40 -c 45 -f -o gpurun_out/r02_full_$c \
    python tools/ncu_target.py $c 3 >> gpurun_out/r02_ncu_$c.log 2>&1
  echo "### exit $?"
done
ls -la gpurun_out/*.ncu-rep
