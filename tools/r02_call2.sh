#!/usr/bin/env bash
# Round 2, call 2 (1 GPU): new bench line at N=1, overlap probes, GPU tests touched by the lease pool.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 300 python -m pytest tests/test_gpu_solver.py tests/test_gpu_outputs.py tests/test_gpu_parity.py -x -q -m gpu
run 400 python bench.py --steps 20 --warmup 5
run 200 python tools/r02_set_probe.py robot_arm POCKIT_B200_GRAPH_PRIORITY=1,0
run 200 python tools/r02_set_probe.py humanoid POCKIT_B200_GRAPH_PRIORITY=1,0
run 120 python tools/timeline.py humanoid
