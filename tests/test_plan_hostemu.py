"""Planner + CUDA-C emitter + job records, checked on the CPU: the generated
per-node programs are compiled as C++ and the job records interpreted with NumPy
(tests/hostemu.py), then compared with the golden vectors of the real reference
and with the oracle."""
import numpy as np
import pytest

from helpers import assert_close, build, golden_cases, load
from hostemu import HostEmu
from pockit_b200 import plan as P

CASES = sorted(golden_cases())


@pytest.fixture(scope="module", params=CASES)
def case(request):
    S = build(request.param)
    return request.param, S, HostEmu(S), load(request.param)


def test_values_match_reference(case):
    name, S, E, g = case
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    assert_close(E.run(P.OBJ, x)[0], g["objective"], "objective")
    assert_close(E.run(P.CONS, x), g["constraints"], "constraints")
    assert_close(E.run(P.GRAD, x), g["gradient"], "gradient")
    assert_close(E.run(P.JAC, x), g["jacobian"], "jacobian")
    assert_close(E.run(P.HESS, x, lam, sigma), g["hessian"], "hessian")
    n_o = len(g["hessian_o"])
    assert_close(E.run(P.HESS, x, np.zeros_like(lam), 1.0)[:n_o], g["hessian_o"], "hessian_o")


@pytest.mark.parametrize("name", ["general_lgl", "robot_arm_lgr_6x20", "rocket_lgl_4x5", "quadrotor_lgr_5x3", "tiny_lgl_2x2"])
def test_fused_expansion_variant_matches_reference(name):
    """DevicePlan(fused=True): the per-node program writes the expanded slots itself."""
    S, g = build(name), load(name)
    E = HostEmu(S, fused=True)
    assert all(len(E.fin[m]["jobs"][P.ST_EXPAND]) == 0 for m in (P.JAC, P.HESS))
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    assert_close(E.run(P.JAC, x), g["jacobian"], "jacobian")
    assert_close(E.run(P.HESS, x, lam, sigma), g["hessian"], "hessian")


@pytest.mark.parametrize("scheme,kw", [("lobatto", dict(mesh=1, num_point=2)), ("radau", dict(mesh=1, num_point=1))])
def test_phase_without_middle_nodes_lowers_and_evaluates(scheme, kw):
    """One interval of minimal order: no middle nodes, so every middle list is empty.  The reference is
    not a parity target here -- its structures and values disagree in length on this mesh and
    ``gradient`` raises (``easyderiv.py:333`` reads element 0 of empty arrays) -- but the lowering must
    not fail, and values, structures and the plan must agree with each other."""
    import importlib

    from pockit_b200 import plan as P
    from pockit_b200 import problems

    S = problems.general(importlib.import_module(f"pockit_b200.{scheme}"), **kw)
    x, lam, sigma = problems.evaluation_point(S, seed=3)
    E = HostEmu(S)
    jr, jc = S.jacobianstructure()
    hr, hc = S.hessianstructure()
    J, H = E.run(P.JAC, x), E.run(P.HESS, x, lam, sigma)
    assert len(J) == len(jr) == len(jc) and len(H) == len(hr) == len(hc)
    assert np.all(np.isfinite(J)) and np.all(np.isfinite(H)) and np.all(hr >= hc)
    assert len(E.run(P.GRAD, x)) == S.L and len(E.run(P.CONS, x)) == len(S.c_lb)
    # the Jacobian is still the derivative of the constraints
    d = np.random.default_rng(0).normal(size=S.L)
    eps = 1e-6
    fd = (E.run(P.CONS, x + eps * d) - E.run(P.CONS, x - eps * d)) / (2 * eps)
    jd = np.zeros(len(S.c_lb))
    np.add.at(jd, jr, J * d[jc])
    np.testing.assert_allclose(jd, fd, rtol=1e-5, atol=1e-6)
