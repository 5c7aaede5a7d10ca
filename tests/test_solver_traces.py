"""Solver-level parity beyond the LQR: SciPy trust-constr wired like the reference adapter
(pockit/optimizer/scipy.py:63-92) must walk the path the REFERENCE walked with its own adapter
(goldens: tests/golden/make_solver_golden.py -- start vector, iteration / evaluation counts and the
objective value at every evaluation) on robot_arm (LGR), the two-phase rocket (FUNC-linked boundary
conditions, static parameters) and the quadrotor (path constraints).

CPU tier: the planner + generated programs (host-emulated, tests/hostemu.py) stand in for the
engine; GPU tier: the CUDA engine through the C-ABI.  20 trust-constr iterations each."""
import importlib

import numpy as np
import pytest

from helpers import GOLDEN
from solver_scipy import solve

CASES = {
    "solver_robot_arm_lgr_4x5": ("robot_arm", "radau", dict(mesh=4, num_point=5)),
    "solver_rocket_lgl_3x4": ("rocket", "lobatto", dict(mesh=3, num_point=4)),
    "solver_quadrotor_lgl_4x4": ("quadrotor", "lobatto", dict(mesh=4, num_point=4)),
}


class EmulatedSystem:
    """The seven callbacks on the host-emulated plan (test infrastructure); everything else is the System's."""

    def __init__(self, S):
        from hostemu import HostEmu
        from pockit_b200 import plan as P

        self._S, self._E, self._P = S, HostEmu(S), P

    def __getattr__(self, k):
        return getattr(self._S, k)

    def objective(self, x):
        return float(self._E.run(self._P.OBJ, x)[0])

    def gradient(self, x):
        return self._E.run(self._P.GRAD, x)

    def constraints(self, x):
        return self._E.run(self._P.CONS, x)

    def jacobian(self, x):
        return self._E.run(self._P.JAC, x)

    def hessian_o(self, x):
        lo = self._S.lowering
        return self._E.run(self._P.HESS, x, np.zeros(lo.m), 1.0)[: lo.nnz_hess_o]

    def hessian_c(self, x, fct_c):
        lo = self._S.lowering
        return self._E.run(self._P.HESS, x, fct_c, 0.0)[lo.nnz_hess_o :]


def _walk(system, g):
    trace = []

    def objective(x):
        trace.append(float(system.objective(x)))
        return trace[-1]

    res = solve(system, g["x0"], options={"maxiter": 20}, objective=objective)
    assert res.nit == int(g["nit"]) and res.nfev == int(g["nfev"]) and res.status == int(g["status"])
    assert len(trace) == len(g["objective_trace"])
    np.testing.assert_allclose(trace, g["objective_trace"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(res.fun, float(g["fun"]), rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(res.constr_violation, float(g["constr_violation"]), rtol=1e-6, atol=1e-10)


@pytest.mark.parametrize("name", sorted(CASES))
def test_trust_constr_trace_matches_reference_on_the_emulated_plan(name):
    from pockit_b200 import problems

    builder, scheme, kw = CASES[name]
    S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
    if builder == "quadrotor":  # the golden was produced with fastmath=True on the reference (Numba); strict here
        S = problems.quadrotor(importlib.import_module(f"pockit_b200.{scheme}"), fastmath=False, **kw)
    _walk(EmulatedSystem(S), np.load(GOLDEN / f"{name}.npz"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_trust_constr_trace_matches_reference_on_the_gpu(name):
    from pockit_b200 import problems

    builder, scheme, kw = CASES[name]
    S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
    _walk(S, np.load(GOLDEN / f"{name}.npz"))
