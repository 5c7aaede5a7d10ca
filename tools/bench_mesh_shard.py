#!/usr/bin/env python
"""One fine mesh sharded over the GPUs of a box (pockit_b200.meshshard): end-to-end eval-sets/s
with host buffers for BASELINE configs[1] (robot_arm LGR 2000x20), checked against the unsharded
engine on rank 0.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_mesh_shard.py [--steps K]

Every rank copies its share of the Jacobian / Hessian values over its own PCIe link into a shared,
page-locked host mapping, so the copy-bound end-to-end rate scales with the number of GPUs.
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--mesh", type=int, default=2000)
    ap.add_argument("--num-point", type=int, default=20)
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
    import pockit_b200.radau as rad
    from pockit_b200 import problems
    from pockit_b200.meshshard import MeshShardedSystem

    S = problems.robot_arm(rad, mesh=args.mesh, num_point=args.num_point)
    x, lam, sigma = problems.evaluation_point(S)
    want = None
    if rank == 0:  # the unsharded engine on the same GPU gives the expected values
        S.pinned_outputs = False
        want = dict(jacobian=S.jacobian(x), hessian=S.hessian(x, lam, sigma), constraints=S.constraints(x),
                    gradient=S.gradient(x), objective=S.objective(x))
        S._engine.close()
        S._engine = None
    ms = MeshShardedSystem(S, device=local)
    ms.pinned_outputs = True
    if rank != 0:
        ms.serve()
    else:
        r = ms.evaluate(x, lam, sigma)
        exact = all(np.array_equal(np.asarray(r[k]), np.asarray(want[k])) for k in want)
        for _ in range(args.warmup):
            ms.evaluate(x, lam, sigma)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ms.evaluate(x, lam, sigma)
        dt = (time.perf_counter() - t0) / args.steps
        each = {}
        for name, fn in (("jacobian", lambda: ms.jacobian(x)), ("hessian", lambda: ms.hessian(x, lam, sigma))):
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            each[name] = 1000.0 * (time.perf_counter() - t0) / 10
        lo = S.lowering
        nj, nh = lo.nnz_jac, lo.nnz_hess_o + lo.nnz_hess_c
        print(json.dumps({
            "tool": "bench_mesh_shard", "workload": f"robot_arm LGR {args.mesh}x{args.num_point}, ONE instance sharded over {world} GPU(s)",
            "n_gpus": world, "steps": args.steps, "e2e_eval_sets_per_s": 1.0 / dt, "e2e_ms_per_step": 1000.0 * dt,
            "ms_per_callback": each, "d2h_bytes_per_step_total": 8 * (1 + lo.r_s + lo.m + nj + nh),
            "bit_identical_to_unsharded": bool(exact), "collective_in_data_path": "none (shared page-locked host mapping)",
        }), flush=True)
        ms.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
