"""Solver-level golden: the reference solved with its own SciPy adapter
(pockit/optimizer/scipy.py:32-100, trust-constr) on the LQR model (BASELINE configs[0]).
Run in the build container only:  python tests/golden/make_solver_golden.py"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference")

import pockit.lobatto as ref
from pockit.optimizer import scipy as ref_scipy
from pockit.optimizer._common import _preprocess
from pockit_b200 import problems

S = problems.lqr(ref, 10, 10)
guess = [ref.constant_guess(S.p[0], 0.0), np.array([0.0])]
x0, _, _ = _preprocess(S, guess, None)
trace = []
orig = S.objective
S.objective = lambda x: (trace.append(float(orig(x.copy()))), trace[-1])[1]
_, res = ref_scipy.solve(S, guess)
np.savez_compressed(
    HERE / "solver_lqr_lgl_10x10.npz", x0=x0, x=res.x, fun=np.float64(res.fun), nit=np.int64(res.nit),
    nfev=np.int64(res.nfev), status=np.int64(res.status), objective_trace=np.array(trace),
)
print("nit", res.nit, "fun", repr(float(res.fun)), "nfev", res.nfev, "status", res.status, "trace", len(trace))
