#!/usr/bin/env bash
# Multi-GPU run of bench.py the way the driver launches it:  bash tools/r02_multi.sh N [steps]
# Every command under its own timeout (a hung rank must not eat the round's GPU budget).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
STEPS=${2:-20}
PORT=$((29500 + N))
echo "### nvidia-smi topo"; timeout 30 nvidia-smi topo -m | head -20
echo "### bench N=$N"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$PORT" \
  bench.py --gpus "$N" --steps "$STEPS" --warmup 5 > "gpurun_out/r02_bench_n$N.json" 2> "gpurun_out/r02_bench_n$N.err"
echo "### exit $?"
tail -c 3000 "gpurun_out/r02_bench_n$N.err"
cat "gpurun_out/r02_bench_n$N.json"
