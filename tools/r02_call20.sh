#!/usr/bin/env bash
# Round 2, call 20 (1 GPU): compute-sanitizer memcheck + racecheck over small parity cases (every kernel family).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 600 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "### $tool exit $?"
  tail -14 gpurun_out/r02_sanitizer_$tool.log
done
