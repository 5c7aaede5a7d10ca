#!/usr/bin/env python
"""BASELINE.json configs[4]: 8192 independent planar_quadrotor instances (LGL 14x6, FIXED initial
state drawn start + U(-0.2, 0.2), rng seed 0) sharded by instance over the ranks, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/bench_batched.py [--instances 8192] [--steps 20]

Total work is fixed (strong scaling).  No data-path collective: every rank evaluates its shard
(device-resident, CUDA events, max over ranks); the objective values are then gathered with
NCCL all_gather, untimed, and checked against rank-local recomputation.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as graft

    graft.build()
    import pockit_b200.lobatto as lob
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.batched import fixed_index, fixed_table
    from pockit_b200.sharding import ShardedBatch

    S = problems.quadrotor(lob)
    n = args.instances
    rng = np.random.default_rng(0)
    fixed = fixed_table(S, n)
    fixed[:, fixed_index(S, 0, "x0", 0)] += rng.uniform(-0.2, 0.2, n)
    fixed[:, fixed_index(S, 0, "x0", 1)] += rng.uniform(-0.2, 0.2, n)
    x0, lam0, sigma = problems.evaluation_point(S)
    X = x0[None, :] + 1e-2 * rng.normal(size=(n, len(x0)))
    LAM = lam0[None, :] + 0.1 * rng.normal(size=(n, len(lam0)))
    sb = ShardedBatch(S, fixed)
    eng = sb.local.engine
    modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    eng.upload(sb._take(X), sb._take(LAM), np.full(len(sb.idx), sigma))
    eng.time_steps(modes, max(3, args.warmup), True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = eng.time_steps(modes, args.steps, True)
    t = torch.tensor([sum(ms) / 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    obj_local = eng.download(P.OBJ).reshape(-1)
    obj = sb.gather(obj_local).reshape(-1)  # NCCL all_gather of the batched results
    ok = bool(np.array_equal(obj[sb.idx], obj_local) and np.all(np.isfinite(obj)))
    if rank == 0:
        lo = S.lowering
        per = 8 * (6 * lo.r_s + 2 * lo.m + lo.nnz_jac + lo.nnz_hess_o + lo.nnz_hess_c)
        step = float(t.item()) / args.steps
        print(json.dumps({
            "config": "planar_quadrotor LGL 14x6, %d instances sharded by instance" % n, "n_gpus": world,
            "scaling": "strong", "instance_sets_per_s": n / step, "ms_per_batch_set": 1e3 * step,
            "algorithmic_GBps_total": per * n / step / 1e9, "gather_ok": ok, "steps": args.steps,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
