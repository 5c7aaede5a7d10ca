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


def test_scipy_adapter_with_x_keyed_cache_walks_the_reference_path():
    """pockit_b200.optimizer.scipy.solve (the reference adapter's wiring + the x-keyed cache): same
    iterations / evaluations / optimum as the reference's own run, with x uploaded once per point and
    fewer engine calls than callbacks."""
    import pockit_b200.lobatto as lob
    from pockit_b200 import problems
    from pockit_b200.guess import Variable
    from pockit_b200.optimizer import scipy as adapter

    g = np.load(GOLDEN / "solver_lqr_lgl_10x10.npz")
    S = problems.lqr(lob, 10, 10)
    x0 = g["x0"]
    guess = [Variable(S.p[0], x0[S.l_p[0] : S.r_p[0]].copy())] + ([x0[S.l_s : S.r_s].copy()] if S.n_s else [])
    guess = guess[0] if len(guess) == 1 else guess
    uploads0 = S.engine.x_uploads
    result, res = adapter.solve(S, guess)
    assert res.nit == int(g["nit"]) and res.nfev == int(g["nfev"]) and res.status == int(g["status"])
    np.testing.assert_allclose(res.fun, float(g["fun"]), rtol=1e-10, atol=0)
    var = result[0] if isinstance(result, list) else result
    sol = np.concatenate([var.data] + ([np.asarray(result[-1])] if S.n_s else []))
    np.testing.assert_allclose(sol, g["x"], rtol=1e-7, atol=1e-9)
    st = res.cache_stats
    assert S.engine.x_uploads - uploads0 == st["points"]           # one host-to-device copy of x per point
    assert st["hits"] > 0 and st["engine_calls"] < st["engine_calls"] + st["hits"]
    # without the cache the solver takes the same path (the cache only removes redundant work)
    _, res2 = adapter.solve(S, guess, cache=False)
    assert res2.nit == res.nit and res2.fun == res.fun
