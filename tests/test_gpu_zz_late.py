"""GPU tests of the outer layers -- continuous error check, cubin cache, mesh sharding (single device and
two processes) -- that sort last on purpose: a surprise here cannot hide the parity tests behind
``pytest -x``.  (Written late in round 1; all of them have run on B200s since.)"""
import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scheme", ["lobatto", "radau"])
def test_check_continuous_known_answers(scheme):
    """The reference's own known-answer test (tests/test_labatto/test_check_lobatto.py:22-36 and the
    radau twin): polynomial trajectories that satisfy x' = u exactly pass, perturbed ones fail."""
    import importlib

    from pockit_b200 import problems

    mod = importlib.import_module(f"pockit_b200.{scheme}")
    S, cases = problems.check_system(mod)
    for value, expected in cases:
        assert S.check_continuous(value) is expected
    with pytest.raises(NotImplementedError):
        S.check_discontinuous(cases[0][0])
    with pytest.raises(ValueError, match="len\\(value\\)"):
        S.check_continuous([cases[0][0][0]])


def test_remeshing_hits_the_cubin_cache():
    """set_discretization + System.update() builds a new engine whose generated programs are the same
    text: the library's process-wide NVRTC cache serves them, and the values follow the new mesh."""
    import pockit_b200.radau as rad
    from oracle.pockit_oracle import OracleSystem
    from pockit_b200 import problems
    from pockit_b200.engine import cubin_cache_stats

    S = problems.robot_arm(rad, mesh=10, num_point=6)
    x, lam, sigma = problems.evaluation_point(S, seed=2)
    S.jacobian(x), S.hessian(x, lam, sigma)
    h0, m0 = cubin_cache_stats()
    S.p[0].set_discretization(14, 9)
    S.update()
    x, lam, sigma = problems.evaluation_point(S, seed=2)
    O = OracleSystem(S)
    assert_close(S.jacobian(x), O.jacobian(x), "jacobian after re-meshing")
    assert_close(S.hessian(x, lam, sigma), O.hessian(x, lam, sigma), "hessian after re-meshing")
    h1, m1 = cubin_cache_stats()
    assert h1 - h0 >= 2 and m1 == m0


# ---------------------------------------------------------------------------------------------
# mesh sharding: the shard plans (tile-range ownership) and the page-locked x path changed after the
# last GPU run of the round; the plans are verified on the host-emulated tier (test_mesh_shard_plan.py)
@pytest.mark.parametrize("G", [2, 5])
def test_mesh_shards_on_one_device_reassemble(G):
    """Mesh sharding: G engines, each planned as rank g of G, write their runs of the
    Jacobian / Hessian into ONE host buffer (pk_engine_set_output_runs); the result must equal the
    unsharded engine bit for bit.  (All shards on cuda:0 here; tools/bench_mesh_shard.py runs one
    rank per GPU.)"""
    import pockit_b200.radau as rad
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.engine import Engine

    S = problems.robot_arm(rad, mesh=150, num_point=12)
    x, lam, sigma = problems.evaluation_point(S, seed=9)
    want_j, want_h = S.jacobian(x), S.hessian(x, lam, sigma)
    jac = np.full(len(want_j), np.nan)
    hess = np.full(len(want_h), np.nan)
    owned = 0
    for g in range(G):
        eng = Engine(S.lowering, shard=(g, G))
        eng.evaluate(x, lam, sigma, modes=[P.JAC, P.HESS], outs=[jac, hess])
        owned += int(eng.fin[P.JAC]["runs"][:, 1].sum())
        if g > 0:
            with pytest.raises(RuntimeError):
                eng.objective(x)
        eng.close()
    assert owned == len(want_j)
    assert np.array_equal(jac, want_j) and np.array_equal(hess, want_h)


def test_weighted_mesh_shards_on_one_device_reassemble():
    """shard = (rank, world, weights): shares sized by the ranks' link rates (meshshard measures them) still
    partition the output; an LGL mesh, so every shard runs the TMA bulk-store expansion."""
    import pockit_b200.lobatto as lob
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.engine import Engine

    S = problems.rocket(lob, mesh=300, num_point=10)
    x, lam, sigma = problems.evaluation_point(S, seed=4)
    want_j, want_h = S.jacobian(x), S.hessian(x, lam, sigma)
    jac, hess = np.full(len(want_j), np.nan), np.full(len(want_h), np.nan)
    w = [13.0, 12.9, 22.5, 25.0]
    owned = []
    for g in range(4):
        eng = Engine(S.lowering, shard=(g, 4, w))
        assert eng.expand_kernel(P.JAC) in ("pk_expand_bulk", "pk_expand_blocks")
        eng.evaluate(x, lam, sigma, modes=[P.JAC, P.HESS], outs=[jac, hess])
        owned.append(int(eng.fin[P.JAC]["runs"][:, 1].sum()))
        eng.close()
    assert sum(owned) == len(want_j) and owned[3] > owned[0]
    assert np.array_equal(jac, want_j) and np.array_equal(hess, want_h)


def _mesh_rank(rank, world, port, q):
    import os
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import pockit_b200.lobatto as lob
    from pockit_b200 import problems
    from pockit_b200.meshshard import MeshShardedSystem

    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = problems.rocket(lob, mesh=120, num_point=9)
    ms = MeshShardedSystem(S, device=0)  # both ranks share cuda:0 in this test
    if rank != 0:
        ms.serve()
    else:
        ok = True
        for seed in (1, 2):
            x, lam, sigma = problems.evaluation_point(S, seed=seed)
            r = ms.evaluate(x, lam, sigma)
            ok &= np.array_equal(r["jacobian"], S.jacobian(x)) and np.array_equal(r["hessian"], S.hessian(x, lam, sigma))
            ok &= np.array_equal(r["constraints"], S.constraints(x)) and np.array_equal(r["gradient"], S.gradient(x))
            ok &= r["objective"] == S.objective(x)
            ok &= np.array_equal(ms.jacobian(x), r["jacobian"])
            ok &= np.array_equal(ms.hessian(x, lam, sigma), r["hessian"])
        q.put(bool(ok))
        ms.close()
    dist.barrier()
    dist.destroy_process_group()


def test_mesh_sharded_system_two_processes():
    """pockit_b200.meshshard end to end: two processes (gloo rendezvous), real engines, results
    assembled in the page-locked shared mapping."""
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mesh_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
