#!/usr/bin/env bash
# First GPU call of the next round: everything that was written after the round-1 GPU budget was
# spent, each step under its own `timeout` (an 8-GPU command that waits for 10 minutes costs 80
# GPU-minutes -- round 1 ended that way).  One GPU, ~4 minutes in total.
#   gpurun --timeout 600 -- 'bash tools/next_round_ab.sh > gpurun_out/next_round_ab.log 2>&1'
set -u
cd "$(dirname "$0")/.."
run() { echo "### $*"; timeout 150 "$@" || echo "### exit $? (timeout 150 s): $*"; }
# 1. late GPU tests (known-answer continuous check, cubin cache) and the whole GPU tier
run python -m pytest tests -m gpu -x -q
# 2. expansion kernels: persistent / parameter-driven / TMA bulk-store -- bit-identical? faster?
run python tools/expand_ab.py robot_arm humanoid rocket
# 3. expression groups in the per-node programs
for c in robot_arm humanoid rocket; do
  run python tools/set_ab.py $c POCKIT_B200_NODE_GROUPS=1,2,4,8
done
# 4. the bulk-store kernel at set level
run python tools/set_ab.py robot_arm POCKIT_B200_EXPAND=params,bulk
# 5. the bench line of the current defaults
run python bench.py
