#!/usr/bin/env bash
# Round 2, first GPU call (1 GPU): the BASELINE-size parity tests, the whole GPU tier, the A/B runs
# that round 1 could not execute, host topology, the bench line.  Every step under its own timeout.
#   gpurun --timeout 1200 -- 'bash tools/r02_call1.sh > gpurun_out/r02_call1.log 2>&1'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 60 nvidia-smi topo -m
run 20 bash -c 'lscpu | head -30; numactl -H 2>/dev/null | head -20; free -g | head -3; nproc'
run 400 python -m pytest tests/test_gpu_baseline_sizes.py -q -m gpu
run 400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_baseline_sizes.py
run 200 python tools/expand_ab.py robot_arm humanoid rocket
for c in robot_arm humanoid; do
  run 150 python tools/set_ab.py $c POCKIT_B200_NODE_GROUPS=1,2,4
done
run 150 python tools/set_ab.py robot_arm POCKIT_B200_EXPAND=params,bulk
run 200 python bench.py --steps 20
