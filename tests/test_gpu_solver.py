"""Solver-level parity on the GPU: SciPy trust-constr driven by the engine-backed system must
walk the same path as the reference solved with its own adapter (golden: tests/golden/
make_solver_golden.py) -- same iteration count, same evaluation count, same optimum."""
import numpy as np
import pytest

from helpers import GOLDEN

pytestmark = pytest.mark.gpu


def test_lqr_trust_constr_matches_reference_run():
    import pockit_b200.lobatto as lob
    from pockit_b200 import problems
    from solver_scipy import solve

    g = np.load(GOLDEN / "solver_lqr_lgl_10x10.npz")
    S = problems.lqr(lob, 10, 10)
    trace = []

    def objective(x):
        trace.append(float(S.objective(x)))
        return trace[-1]

    res = solve(S, g["x0"], objective=objective)
    assert res.nit == int(g["nit"]) and res.nfev == int(g["nfev"]) and res.status == int(g["status"])
    np.testing.assert_allclose(res.fun, float(g["fun"]), rtol=1e-10, atol=0)
    np.testing.assert_allclose(trace, g["objective_trace"], rtol=1e-9, atol=1e-12)
    # boundary slots are not optimisation variables (the reference overwrites them in place,
    # phasebase.py:840-847; pockit/optimizer/_common.py:39-63 re-applies them) -- substitute first
    x, p, s_ = res.x.copy(), S.p[0], res.x[S.l_s : S.r_s]
    for i in range(p.n_x):
        x[p.l_v[i]] = p._value_boundary_condition(p.info_bc_0[i], x[p.l_v[i]], s_)
        x[p.r_v[i] - 1] = p._value_boundary_condition(p.info_bc_f[i], x[p.r_v[i] - 1], s_)
    x[p.L - 2] = p._value_boundary_condition(p.info_t_0, x[p.L - 2], s_)
    x[p.L - 1] = p._value_boundary_condition(p.info_t_f, x[p.L - 1], s_)
    np.testing.assert_allclose(x, g["x"], rtol=1e-7, atol=1e-9)
