#!/usr/bin/env python
"""Target for compute-sanitizer runs: one small case of every kernel family -- the
parameter-driven column walk, its TMA bulk-store variant, the persistent kernel on an hp mesh, the batch
kernel, table-driven pieces, defects / reductions / gradient gathers, weighted mesh shards -- each checked
against the oracle.

    compute-sanitizer --tool memcheck  --error-exitcode 9 python tools/sanitize_target.py
    compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
g.build()
from pockit_b200 import problems, plan as P
from pockit_b200.engine import Engine
from pockit_b200.batched import BatchedSystem, fixed_table
from oracle.pockit_oracle import OracleSystem
def check(S, name, env=None):
    for k, v in (env or {}).items(): os.environ[k] = v
    x, lam, sigma = problems.evaluation_point(S, seed=3)
    O = OracleSystem(S)
    eng = Engine(S.lowering, fastmath=S._fastmath)
    r = eng.evaluate(x, lam, sigma)
    eng.upload(x, lam, sigma); eng.time_steps([P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS], 2, flush_l2=False); eng.sync()
    ok = np.allclose(r[P.JAC], O.jacobian(x), rtol=1e-12, atol=1e-14) and np.allclose(r[P.HESS], O.hessian(x, lam, sigma), rtol=1e-12, atol=1e-14)
    ok &= np.allclose(np.atleast_1d(eng.download(P.GRAD)), O.gradient(x), rtol=1e-12, atol=1e-14) and np.allclose(eng.download(P.CONS), O.constraints(x), rtol=1e-12, atol=1e-14)
    print(name, eng.expand_kernel(P.HESS), "ok" if ok else "MISMATCH", flush=True)
    eng.close()
    for k in (env or {}): os.environ.pop(k, None)
import pockit_b200.radau as rad, pockit_b200.lobatto as lob
ONLY_BATCH = os.environ.get("PK_SAN_ONLY") == "batch"  # just the batch kernels (both thread mappings)
if not ONLY_BATCH:
  check(problems.robot_arm(rad, mesh=64, num_point=20), "robot_arm 64x20 cols")
  check(problems.robot_arm(rad, mesh=64, num_point=20), "robot_arm 64x20 bulk", {"POCKIT_B200_EXPAND": "bulk"})
  check(problems.rocket(lob, mesh=120, num_point=10), "rocket 2x120x10 (bulk by default)")
  check(problems.rocket(lob, mesh=20, num_point=[4] * 6 + [7] * 8 + [3, 5, 5, 5, 8, 8]), "rocket hp mesh (persistent blocks)")
  check(problems.humanoid(lob, mesh=20, num_point=10), "humanoid 20x10")
  check(problems.general(lob), "general lgl")
  check(problems.general(rad), "general lgr")
S = problems.quadrotor(lob, fastmath=False)
B = 64
x0, lam0, _ = problems.evaluation_point(S)
rng = np.random.default_rng(0)
X = x0[None, :] + 1e-2 * rng.normal(size=(B, len(x0))); LAM = np.tile(lam0, (B, 1)); sig = np.ones(B)
O = OracleSystem(S)
for mapping in ("", "batch"):  # default (slot order) and the column mapping
    if mapping: os.environ["POCKIT_B200_EXPAND"] = mapping
    bs = BatchedSystem(S, fixed_table(S, B))
    J = bs.jacobian(X); H = bs.hessian(X, LAM, sig); C = bs.constraints(X); G = bs.gradient(X)
    ok = all(np.allclose(J[b], O.jacobian(X[b]), rtol=1e-12, atol=1e-14) and np.allclose(H[b], O.hessian(X[b], LAM[b], 1.0), rtol=1e-12, atol=1e-14) for b in (0, 17, 63))
    print("quadrotor batch 64", bs.engine.expand_kernel(P.JAC), "ok" if ok else "MISMATCH", flush=True)
    bs.close()
    os.environ.pop("POCKIT_B200_EXPAND", None)
for g_ in range(0 if ONLY_BATCH else 3):
    eng = Engine(problems.robot_arm(rad, mesh=150, num_point=12).lowering, shard=(g_, 3, [1.0, 2.0, 0.7]))
    eng.close()
print("san target done", flush=True)
