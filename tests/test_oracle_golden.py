"""Pin the CPU oracle (oracle/pockit_oracle.py) and the host planner's COO
patterns against golden vectors produced by the real reference."""
import numpy as np
import pytest

from helpers import assert_close, build, golden_cases, load

CASES = sorted(golden_cases())


@pytest.fixture(scope="module", params=CASES)
def case(request):
    from oracle.pockit_oracle import OracleSystem

    S = build(request.param)
    return request.param, S, OracleSystem(S), load(request.param)


def test_layout_and_bounds(case):
    name, S, O, g = case
    assert S.L == len(g["x"]) and len(S.c_lb) == len(g["lam"])
    for k in ("v_lb", "v_ub", "c_lb", "c_ub"):
        assert np.array_equal(getattr(S, k), g[k]), k


def test_structures_bit_exact(case):
    name, S, O, g = case
    for who in (S, O):
        jr, jc = who.jacobianstructure()
        hr, hc = who.hessianstructure()
        assert np.array_equal(jr, g["jac_row"]) and np.array_equal(jc, g["jac_col"])
        assert np.array_equal(hr, g["hess_row"]) and np.array_equal(hc, g["hess_col"])


def test_oracle_values(case):
    name, S, O, g = case
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    assert_close(O.objective(x.copy()), g["objective"], "objective")
    assert_close(O.gradient(x.copy()), g["gradient"], "gradient")
    assert_close(O.constraints(x.copy()), g["constraints"], "constraints")
    assert_close(O.jacobian(x.copy()), g["jacobian"], "jacobian")
    assert_close(O.hessian_o(x.copy()), g["hessian_o"], "hessian_o")
    assert_close(O.hessian(x.copy(), lam, sigma), g["hessian"], "hessian")
