"""N > 1 host logic on CPU: two gloo ranks shard a batch of instances, evaluate their
shards (with the host-emulated plan standing in for the GPU engine) and gather; the
assembled result must equal the unsharded evaluation."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import pockit_b200.lobatto as lob
    from hostemu import HostEmu
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.batched import fixed_index, fixed_table
    from pockit_b200.sharding import ShardedBatch

    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = problems.quadrotor(lob, mesh=3, num_point=4, fastmath=False)
    n = 5
    rng = np.random.default_rng(0)
    fixed = fixed_table(S, n)
    fixed[:, fixed_index(S, 0, "x0", 0)] += rng.uniform(-0.2, 0.2, n)
    fixed[:, fixed_index(S, 0, "x0", 1)] += rng.uniform(-0.2, 0.2, n)
    x0, lam0, _ = problems.evaluation_point(S)
    X = x0[None, :] + 1e-2 * rng.normal(size=(n, len(x0)))
    LAM = lam0[None, :] + 0.1 * rng.normal(size=(n, len(lam0)))

    class Emu:  # CPU stand-in for BatchedSystem (test infrastructure)
        def __init__(self, f):
            self.e = HostEmu(S, batch=len(f), fixed=f)

        def objective(self, X):
            return self.e.run(P.OBJ, X).reshape(-1)

        def jacobian(self, X):
            return self.e.run(P.JAC, X)

        def hessian(self, X, lam, sig):
            return self.e.run(P.HESS, X, lam, sig)

    sb = ShardedBatch(S, fixed, make_evaluator=lambda f: Emu(f))
    obj = sb.gather(sb.objective(X))
    jac = sb.gather(sb.jacobian(X))
    hes = sb.gather(sb.hessian(X, LAM, 0.5))
    if rank == 0:
        full = Emu(fixed)
        ok = (
            np.array_equal(obj.reshape(-1), full.objective(X))
            and np.array_equal(jac, full.jacobian(X).reshape(n, -1))
            and np.array_equal(hes, full.hessian(X, LAM, np.full(n, 0.5)).reshape(n, -1))
            and len(set(np.round(obj.reshape(-1), 12))) == n  # the instances really differ
        )
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_batch_matches_unsharded():
    import torch.multiprocessing as mp

    from pockit_b200.sharding import shard_indices

    assert shard_indices(5, 0, 2).tolist() == [0, 2, 4] and shard_indices(5, 1, 2).tolist() == [1, 3]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
