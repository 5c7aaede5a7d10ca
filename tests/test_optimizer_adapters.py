"""CPU tier: the solver adapters' host logic -- guess packing (reference behaviour and error texts,
pockit/optimizer/_common.py:9-63) and the x-keyed evaluation cache, with the host-emulated plan
standing in for the CUDA engine."""
import numpy as np
import pytest

from helpers import build, load
from hostemu import HostEmu
from pockit_b200 import plan as P
from pockit_b200.optimizer._cache import CachedCallbacks
from pockit_b200.optimizer._common import pack_guess, unpack_solution


class EmuEngine:
    """Engine.evaluate() stand-in (test infrastructure): records what was asked for."""

    def __init__(self, S):
        self.e = HostEmu(S)
        self.lowering = S.lowering
        self.compacted = set()
        self.calls = []
        self.x = None
        self.x_uploads = 0

    def upload(self, x):  # any other entry point of the engine that replaces the resident x
        self.x = np.array(x, copy=True)
        self.x_uploads += 1

    def evaluate(self, x, fct_c=None, fct_o=None, modes=None, outs=None):
        self.calls.append((x is not None, tuple(modes)))
        if x is not None:
            self.upload(x)
        res = {}
        for m in modes:
            v = self.e.run(m, self.x, fct_c, None if fct_o is None else float(np.asarray(fct_o).reshape(-1)[0]))
            res[m] = v[0] if m == P.OBJ else v
        return res


    def hessian_o(self, x):
        self.calls.append((x is not None, ("hessian_o",)))
        if x is not None:
            self.upload(x)
        return self.e.run(P.HESS, self.x, np.zeros(self.lowering.m), 1.0)[: self.lowering.nnz_hess_o]

    def hessian_c(self, x, fct_c):
        self.calls.append((x is not None, ("hessian_c",)))
        if x is not None:
            self.upload(x)
        return self.e.run(P.HESS, self.x, fct_c, 0.0)[self.lowering.nnz_hess_o:]


class FakeSystem:
    def __init__(self, S):
        self._S = S
        self.engine = EmuEngine(S)

    def __getattr__(self, k):
        return getattr(self._S, k)


def test_cache_groups_callbacks_and_uploads_once_per_point():
    S, g = build("robot_arm_lgr_6x20"), load("robot_arm_lgr_6x20")
    F = FakeSystem(S)
    cb = CachedCallbacks(F)
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    f = cb.objective(x)
    c = cb.constraints(x.copy())          # same point, different array object: served from the cache
    gr = cb.gradient(x)
    J = cb.jacobian(x)
    H = cb.hessian(x, lam, sigma)
    np.testing.assert_allclose(f, g["objective"], rtol=1e-12)
    np.testing.assert_allclose(c, g["constraints"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(gr, g["gradient"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(J, g["jacobian"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(H, g["hessian"], rtol=1e-12, atol=1e-14)
    # three engine calls for five callbacks, x sent with the first one only
    assert F.engine.calls == [(True, (P.OBJ, P.CONS)), (False, (P.GRAD, P.JAC)), (False, (P.HESS,))]
    assert cb.stats == {"points": 1, "engine_calls": 3, "hits": 2}
    # a line-search trial point: new upload, only the {f, g} group
    x2 = x + 1e-3
    cb.objective(x2); cb.constraints(x2)
    assert F.engine.calls[-1] == (True, (P.OBJ, P.CONS)) and len(F.engine.calls) == 4
    # back to the first point: the cache holds one point, so it is evaluated again
    cb.jacobian(x)
    assert F.engine.calls[-1] == (True, (P.GRAD, P.JAC))
    # SciPy's split Hessians
    n_o = S.lowering.nnz_hess_o
    np.testing.assert_allclose(cb.hessian_o(x), g["hessian_o"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(cb.hessian_c(x, lam), g["hessian"][n_o:], rtol=1e-12, atol=1e-14)
    # attribute passthrough (bounds / structures are the system's own)
    assert cb.L == S.L and np.array_equal(cb.jacobianstructure()[0], S.jacobianstructure()[0])


def test_cache_notices_a_foreign_upload():
    """The device copy of x belongs to the engine, not to the cache: when anything else uploads a
    point in between (System.objective, check_continuous, Engine.upload ...), the next cached
    callback at the old point must send x again instead of evaluating at the foreign one."""
    S, g = build("robot_arm_lgr_6x20"), load("robot_arm_lgr_6x20")
    F = FakeSystem(S)
    cb = CachedCallbacks(F)
    x = g["x"]
    cb.objective(x)
    F.engine.upload(x + 0.25)  # e.g. an `intermediate` hook calling system.check_continuous(x_k)
    J = cb.jacobian(x)
    assert F.engine.calls[-1] == (True, (P.GRAD, P.JAC))
    np.testing.assert_allclose(J, g["jacobian"], rtol=1e-12, atol=1e-14)
    H = cb.hessian(x, g["lam"], float(g["sigma"]))  # nothing in between: resident again
    assert F.engine.calls[-1] == (False, (P.HESS,))
    np.testing.assert_allclose(H, g["hessian"], rtol=1e-12, atol=1e-14)


def test_cache_without_grouping_and_nan_points():
    S, g = build("lqr_lgl_10x10"), load("lqr_lgl_10x10")
    F = FakeSystem(S)
    cb = CachedCallbacks(F, grouping=False)
    x = g["x"]
    cb.objective(x); cb.constraints(x)
    assert F.engine.calls == [(True, (P.OBJ,)), (False, (P.CONS,))]
    xn = x.copy(); xn[3] = np.nan
    cb.objective(xn); cb.objective(xn)     # NaN never equals itself: evaluated twice, never served stale
    assert [c[0] for c in F.engine.calls[-2:]] == [True, True]


def test_guess_packing_follows_the_reference():
    import pockit_b200.lobatto as lob
    from pockit_b200 import problems
    from pockit_b200.guess import linear_guess

    S = problems.rocket(lob, mesh=4, num_point=5)
    guesses = [linear_guess(p, 0.5) for p in S.p]
    s0 = np.linspace(1.0, 2.0, S.n_s)
    x0, single, opts = pack_guess(S, guesses + [s0], None)
    assert not single and opts == {} and len(x0) == S.L
    for i, gq in enumerate(guesses):
        assert np.array_equal(x0[S.l_p[i] : S.r_p[i]], gq.data)
    assert np.array_equal(x0[S.l_s : S.r_s], s0)
    with pytest.raises(ValueError, match="number of phases \\+ 1"):
        pack_guess(S, guesses, None)
    back = unpack_solution(S, x0, False)
    assert len(back) == S.n_p + 1 and np.array_equal(back[-1], s0)
    # boundary slots come back substituted (FIXED / FUNC values), everything else untouched
    p = S.p[0]
    for j in range(p.n_x):
        assert back[0].data[p.l_v[j]] == p._value_boundary_condition(p.info_bc_0[j], x0[S.l_p[0] + p.l_v[j]], s0)
    L = problems.lqr(lob, 3, 3)
    gq = linear_guess(L.p[0], 0.5) if not L.n_s else None
    if gq is not None:
        x0, single, _ = pack_guess(L, gq, {"maxiter": 3})
        assert single and isinstance(unpack_solution(L, x0, True), type(gq))


def test_ipopt_adapter_needs_cyipopt():
    import importlib.util

    import pockit_b200.lobatto as lob
    from pockit_b200 import problems
    from pockit_b200.optimizer import ipopt

    if importlib.util.find_spec("cyipopt") is not None:
        pytest.skip("cyipopt is installed here")
    with pytest.raises(ImportError, match="cyipopt"):
        ipopt.solve(problems.lqr(lob, 3, 3), None)


def test_ipopt_adapter_wiring_with_a_stand_in_for_cyipopt(monkeypatch):
    """cyipopt / libipopt are not in this image, so Ipopt itself cannot run; the adapter's own code can: a
    stand-in module with cyipopt's interface (``Problem(n, m, problem_obj, lb, ub, cl, cu)``,
    ``add_option``, ``solve(x0) -> (x, info)``) drives the callbacks in Ipopt's order and with Ipopt's
    argument convention ``hessian(x, lagrange, obj_factor)`` (``pockit/optimizer/ipopt.py:41-53``), on the
    host-emulated plan.  Checks sizes, structures, the x-keyed cache and the returned solution layout."""
    import sys
    import types

    S, g = build("robot_arm_lgr_6x20"), load("robot_arm_lgr_6x20")
    F = FakeSystem(S)
    seen = {}

    class Problem:
        def __init__(self, n, m, problem_obj, lb, ub, cl, cu):
            assert n == S.L and m == len(S.c_lb) and len(lb) == len(ub) == n and len(cl) == len(cu) == m
            self.n, self.m, self.obj, self.options = n, m, problem_obj, {}

        def add_option(self, k, v):
            self.options[k] = v

        def solve(self, x0):
            o = self.obj
            jr, jc = o.jacobianstructure()
            hr, hc = o.hessianstructure()
            x = np.array(g["x"], copy=True)  # evaluate at the golden point, as one Ipopt iteration would
            seen["f"], seen["grad"], seen["c"] = o.objective(x), o.gradient(x), o.constraints(x)
            seen["jac"] = o.jacobian(x)
            seen["hess"] = o.hessian(x, g["lam"], float(g["sigma"]))  # (x, lagrange, obj_factor)
            assert len(seen["jac"]) == len(jr) == len(jc) and len(seen["hess"]) == len(hr) == len(hc)
            return x0, {"status": 0, "obj_val": float(seen["f"]), "options": dict(self.options)}

    monkeypatch.setitem(sys.modules, "cyipopt", types.SimpleNamespace(Problem=Problem))
    from pockit_b200.guess import Variable
    from pockit_b200.optimizer import ipopt

    x0 = g["x"]
    guess = Variable(S.p[0], x0[S.l_p[0] : S.r_p[0]].copy())
    result, info = ipopt.solve(F, guess, {"max_iter": 7, "tol": 1e-9})
    assert info["options"] == {"max_iter": 7, "tol": 1e-9} and info["status"] == 0
    np.testing.assert_allclose(seen["f"], g["objective"], rtol=1e-12)
    np.testing.assert_allclose(seen["grad"], g["gradient"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(seen["c"], g["constraints"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(seen["jac"], g["jacobian"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(seen["hess"], g["hessian"], rtol=1e-12, atol=1e-14)
    # five callbacks at one point: x sent once, three engine calls ({f, g}, {grad f, J}, H)
    assert info["cache_stats"] == {"points": 1, "engine_calls": 3, "hits": 2}
    assert isinstance(result, Variable) and len(result.data) == S.p[0].L


def test_scipy_adapter_solves_the_lqr_like_the_reference_on_the_host_emulated_plan():
    """Solver-level parity without a GPU: pockit_b200.optimizer.scipy.solve wired to the
    host-emulated plan must walk the reference's own trust-constr path on the LQR
    (tests/golden/solver_lqr_lgl_10x10.npz: iterations, evaluations, optimum)."""
    import pockit_b200.lobatto as lob
    from helpers import GOLDEN
    from pockit_b200 import problems
    from pockit_b200.guess import Variable
    from pockit_b200.optimizer import scipy as adapter

    g = np.load(GOLDEN / "solver_lqr_lgl_10x10.npz")
    S = problems.lqr(lob, 10, 10)
    F = FakeSystem(S)
    x0 = g["x0"]
    guess = [Variable(S.p[0], x0[S.l_p[0] : S.r_p[0]].copy())] + ([x0[S.l_s : S.r_s].copy()] if S.n_s else [])
    guess = guess[0] if len(guess) == 1 else guess
    result, res = adapter.solve(F, guess)
    assert res.nit == int(g["nit"]) and res.nfev == int(g["nfev"]) and res.status == int(g["status"])
    np.testing.assert_allclose(res.fun, float(g["fun"]), rtol=1e-10, atol=0)
    var = result[0] if isinstance(result, list) else result
    sol = np.concatenate([var.data] + ([np.asarray(result[-1])] if S.n_s else []))
    np.testing.assert_allclose(sol, g["x"], rtol=1e-7, atol=1e-9)
    st = res.cache_stats
    uploads = sum(1 for sent_x, _ in F.engine.calls if sent_x)
    assert uploads == st["points"] and st["hits"] > 0


def test_tools_parse():
    """The measurement tools are part of the deliverable: they must at least compile."""
    import py_compile
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    for f in sorted((root / "tools").glob("*.py")) + [root / "bench.py", root / "__graft_entry__.py"]:
        py_compile.compile(str(f), doraise=True)


def test_mirrored_hessian_known_answer():
    """The reference's own test of the lower-triangle -> full matrix reflection
    (tests/test_optimizer/test_optimizer_scipy.py:7-15): duplicates summed, diagonal counted once."""
    from pockit_b200.optimizer.scipy import _mirrored

    row = np.array([3, 2, 2, 0, 2, 2, 1])
    col = np.array([3, 2, 0, 0, 1, 2, 0])
    full = _mirrored(lambda _: np.arange(7) ** 2, row, col, 4)(None).toarray()
    assert np.all(full == np.array([[9, 36, 4, 0], [36, 0, 16, 0], [4, 16, 1 + 25, 0], [0, 0, 0, 0]]))
