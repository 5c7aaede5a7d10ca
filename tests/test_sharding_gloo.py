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


# ---------------------------------------------------------------------------------------------
# one fine mesh split over two ranks (pockit_b200.meshshard): shared host buffer, no collective
def _mesh_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import pockit_b200.radau as rad
    from hostemu import HostEmu
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.meshshard import MeshShardedSystem

    P.SPLIT_MIN = 16  # let the small test problem split its table / constant runs too
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = problems.robot_arm(rad, mesh=12, num_point=5)

    class EmuEngine:  # CPU stand-in for Engine(shard=...) (test infrastructure)
        def __init__(self, lowering, shard):
            self.e = HostEmu(S, shard=shard)

        def evaluate(self, x, fct_c=None, fct_o=None, modes=None, outs=None, wait=True):
            if x is not None:  # x travels with the first call of a point; later calls use the resident copy
                self.x = np.array(x, copy=True)
            x = self.x
            res = {}
            for k, m in enumerate(modes):
                full = self.e.run(m, x, fct_c, None if fct_o is None else float(np.asarray(fct_o).reshape(-1)[0]))
                out = outs[k] if outs is not None and outs[k] is not None else np.full(len(full), np.nan)
                for off, cnt in self.e.fin[m]["runs"]:
                    out[off : off + cnt] = full[off : off + cnt]
                res[m] = out[0] if m == P.OBJ else out
            return res

        def sync(self):
            pass

        def objective(self, x):
            return self.e.run(P.OBJ, x)[0]

    # rank 1 finishes its start-up late: rank 0 has published its first point by then (regression:
    # a worker that initialised its sequence number from the live counter waited forever)
    ms = MeshShardedSystem(S, make_engine=EmuEngine, _startup_delay=1.5 if rank == 1 else 0.0)
    if rank != 0:
        ms.serve()
    else:
        whole = HostEmu(S)
        ok = True
        for seed in (3, 4):
            x, lam, sigma = problems.evaluation_point(S, seed=seed)
            ok &= np.array_equal(ms.jacobian(x), whole.run(P.JAC, x))
            ok &= np.array_equal(ms.hessian(x, lam, 0.5), whole.run(P.HESS, x, lam, 0.5))
            r = ms.evaluate(x, lam, sigma)
            ok &= np.array_equal(r["jacobian"], whole.run(P.JAC, x))
            ok &= np.array_equal(r["hessian"], whole.run(P.HESS, x, lam, sigma))
            ok &= np.array_equal(r["constraints"], whole.run(P.CONS, x))
            ok &= np.array_equal(r["gradient"], whole.run(P.GRAD, x))
            ok &= r["objective"] == whole.run(P.OBJ, x)[0]
        # both ranks really contributed: rank 1 owns part of the Jacobian
        ok &= len(HostEmu(S, shard=(1, 2)).fin[P.JAC]["runs"]) > 0
        q.put(bool(ok))
        ms.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_mesh_shard_matches_unsharded():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mesh_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
