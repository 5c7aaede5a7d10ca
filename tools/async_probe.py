#!/usr/bin/env python
"""Host-side cost of the asynchronous set call (pk_eval_set_async): how long does the CALL take, with
which destination memory?   python tools/async_probe.py"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import __graft_entry__ as graft

    graft.build()
    import pockit_b200.radau as rad
    from pockit_b200 import plan as P
    from pockit_b200 import problems
    from pockit_b200.engine import Engine, PinnedArray

    S = problems.robot_arm(rad, 2000, 20)
    x, lam, sigma = problems.evaluation_point(S)
    eng = Engine(S.lowering)
    lo = S.lowering
    pj, ph = PinnedArray(lo.nnz_jac), PinnedArray(lo.nnz_hess_o + lo.nnz_hess_c)
    px = PinnedArray(lo.r_s)
    px.array[:] = x
    eng.evaluate(x, lam, sigma)
    for label, modes, outs, xin in (
        ("small3 + jac, leases, pageable x", [P.OBJ, P.GRAD, P.CONS, P.JAC], None, x),
        ("jac only, pinned out, pageable x", [P.JAC], [pj.array], x),
        ("jac only, pinned out, pinned x", [P.JAC], [pj.array], px.array),
        ("small3 only, leases, pinned x", [P.OBJ, P.GRAD, P.CONS], None, px.array),
        ("grad only, lease, pinned x", [P.GRAD], None, px.array),
        ("obj only, pinned x", [P.OBJ], None, px.array),
    ):
        ts = []
        for _ in range(8):
            eng.sync()
            t0 = time.perf_counter()
            r = eng.evaluate(xin, modes=modes, outs=outs, wait=False)
            t1 = time.perf_counter()
            eng.sync()
            t2 = time.perf_counter()
            ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t0)))
            del r
        ts.sort()
        print(f"{label:40s} call {ts[len(ts)//2][0]:7.3f} ms   call+sync {ts[len(ts)//2][1]:7.3f} ms", flush=True)


def mesh_world1():
    """The sharded path with one rank: the shared /dev/shm mapping (page-locked with cudaHostRegister) as
    source of x and destination of the Jacobian / Hessian values."""
    import pockit_b200.radau as rad
    from pockit_b200 import problems
    from pockit_b200.meshshard import MeshShardedSystem

    S = problems.robot_arm(rad, 2000, 20)
    x, lam, sigma = problems.evaluation_point(S)
    ms = MeshShardedSystem(S, rank=0, world=1, device=0)
    ms.pinned_outputs = True
    for _ in range(4):
        ms.evaluate(x, lam, sigma)
    print("mesh world=1 timeline (ms):", {k: (round(v, 3) if not isinstance(v, list) else v) for k, v in ms.last_timeline.items()}, flush=True)
    t0 = time.perf_counter()
    for _ in range(10):
        ms.evaluate(x, lam, sigma)
    print(f"mesh world=1: {100 * (time.perf_counter() - t0):.3f} ms per set", flush=True)
    ms.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mesh":
        mesh_world1()
        sys.exit(0)
    main()
