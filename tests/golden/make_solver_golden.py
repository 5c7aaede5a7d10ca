"""Solver-level goldens: the reference solved with its own SciPy adapter
(pockit/optimizer/scipy.py:32-100, trust-constr).  Run in the build container only:

    python tests/golden/make_solver_golden.py [case ...]

Every case stores the start vector the adapter packed, the result, iteration / evaluation counts and
the objective value at every evaluation (the trace): an engine-backed system driven through the same
wiring must walk the same path.  LQR (BASELINE configs[0]) runs to convergence; the other models run
a fixed number of trust-constr iterations on a small mesh (the trace is what is compared).
"""
import importlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference")

# name: (builder, scheme, kwargs, guess, optimizer options)
CASES = {
    "solver_lqr_lgl_10x10": ("lqr", "lobatto", dict(mesh=10, num_point=10), ("constant", 0.0), None),
    "solver_robot_arm_lgr_4x5": ("robot_arm", "radau", dict(mesh=4, num_point=5), ("point", 3), {"maxiter": 20}),
    "solver_rocket_lgl_3x4": ("rocket", "lobatto", dict(mesh=3, num_point=4), ("point", 5), {"maxiter": 20}),
    "solver_quadrotor_lgl_4x4": ("quadrotor", "lobatto", dict(mesh=4, num_point=4), ("point", 7), {"maxiter": 20}),
}


def make_guess(mod, S, spec):
    """``('constant', v)``: the reference's constant guess; ``('point', seed)``: the synthetic
    evaluation point of ``pockit_b200.problems`` wrapped in Variables (works on both implementations)."""
    from pockit_b200 import problems

    if spec[0] == "constant":
        guess = [mod.constant_guess(p, spec[1]) for p in S.p]
        if S.n_s:
            guess.append(np.zeros(S.n_s))
    else:
        x, _, _ = problems.evaluation_point(S, seed=spec[1])
        guess = [mod.Variable(p, x[int(S.l_p[i]) : int(S.r_p[i])].copy()) for i, p in enumerate(S.p)]
        if S.n_s:
            guess.append(x[int(S.l_s) : int(S.r_s)].copy())
    return guess[0] if len(guess) == 1 else guess


def main(names):
    from pockit.optimizer import scipy as ref_scipy
    from pockit.optimizer._common import _preprocess
    from pockit_b200 import problems

    for name in names:
        builder, scheme, kw, gspec, opts = CASES[name]
        mod = importlib.import_module(f"pockit.{scheme}")
        S = problems.BUILDERS[builder](mod, **kw)
        guess = make_guess(mod, S, gspec)
        x0, _, _ = _preprocess(S, guess, None)
        trace = []
        orig = S.objective
        S.objective = lambda x: (trace.append(float(orig(x.copy()))), trace[-1])[1]
        _, res = ref_scipy.solve(S, guess, dict(opts) if opts else None)
        np.savez_compressed(
            HERE / f"{name}.npz", x0=x0, x=res.x, fun=np.float64(res.fun), nit=np.int64(res.nit),
            nfev=np.int64(res.nfev), status=np.int64(res.status), objective_trace=np.array(trace),
            constr_violation=np.float64(res.constr_violation),
        )
        print(name, "nit", res.nit, "fun", repr(float(res.fun)), "nfev", res.nfev, "status", res.status, "trace", len(trace), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
