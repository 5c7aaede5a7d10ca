#!/usr/bin/env bash
# Round 2, last 1-GPU run on the final library (slot-order batch kernel as default): whole GPU tier, smoke,
# the bench line, memcheck of the two batch kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local t=$1; shift; echo "### $*"; timeout "$t" "$@"; echo "### exit $? : $*"; }
run 600 python -m pytest tests -m gpu -x -q
run 150 python -c "import __graft_entry__ as g; g.smoke()"
run 400 python bench.py --steps 20 --warmup 5
PK_SAN_ONLY=batch run 70 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py
