#!/usr/bin/env bash
# Round 2, final 1-GPU evidence: whole GPU tier, smoke, the bench line, its ncu launch list, a full capture of one set.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 700 python -m pytest tests -m gpu -x -q
run 200 python -c "import __graft_entry__ as g; g.smoke()"
run 700 python bench.py --steps 20 --warmup 5
echo "### ncu launch list of bench.py --steps 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv \
  python bench.py --steps 2 --warmup 3 --no-all-configs --no-c5 --no-compact --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "### exit $?"
echo "### ncu --set full robot_arm (third evaluation set)"
timeout 600 ncu --set full --clock-control none --import-source on -s 20 -c 10 -f -o gpurun_out/r02_full_robot_arm \
  python tools/ncu_target.py robot_arm 3 > gpurun_out/r02_ncu_robot_arm.log 2>&1
echo "### exit $?"
