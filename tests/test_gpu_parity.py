"""GPU parity tests proper: the CUDA path (through the C-ABI) against the golden
vectors of the real reference and against the CPU oracle on seeded inputs."""
import numpy as np
import pytest

from helpers import assert_close, build, golden_cases, load

pytestmark = pytest.mark.gpu
CASES = sorted(golden_cases())


@pytest.fixture(scope="module", params=CASES)
def case(request):
    S = build(request.param)
    return request.param, S, load(request.param)


def test_callbacks_match_reference_golden(case):
    name, S, g = case
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    assert_close(S.objective(x.copy()), g["objective"], "objective")
    assert_close(S.gradient(x.copy()), g["gradient"], "gradient")
    assert_close(S.constraints(x.copy()), g["constraints"], "constraints")
    assert_close(S.jacobian(x.copy()), g["jacobian"], "jacobian")
    assert_close(S.hessian(x.copy(), lam, sigma), g["hessian"], "hessian")
    assert_close(S.hessian_o(x.copy()), g["hessian_o"], "hessian_o")
    n_o = len(g["hessian_o"])
    assert_close(S.hessian_c(x.copy(), lam), g["hessian"][n_o:], "hessian_c")
    # the engine must not write boundary values into the caller's x
    x2 = x.copy()
    S.jacobian(x2)
    assert np.array_equal(x2, x)


@pytest.mark.parametrize(
    "builder,scheme,kw",
    [
        ("robot_arm", "radau", dict(mesh=40, num_point=20)),
        ("robot_arm", "lobatto", dict(mesh=64, num_point=7)),
        ("rocket", "lobatto", dict(mesh=50, num_point=10)),
        ("quadrotor", "lobatto", dict(mesh=14, num_point=6)),
        ("humanoid", "lobatto", dict(mesh=20, num_point=10)),
        ("lqr", "radau", dict(mesh=[0, 0.1, 0.15, 0.4, 0.7, 1.0], num_point=[4, 7, 3, 9, 5])),
    ],
)
def test_callbacks_match_oracle(builder, scheme, kw):
    import importlib

    from oracle.pockit_oracle import OracleSystem
    from pockit_b200 import problems

    S = problems.BUILDERS[builder](importlib.import_module(f"pockit_b200.{scheme}"), **kw)
    O = OracleSystem(S)
    x, lam, sigma = problems.evaluation_point(S, seed=7)
    assert_close(S.objective(x), O.objective(x), "objective")
    assert_close(S.gradient(x), O.gradient(x), "gradient")
    assert_close(S.constraints(x), O.constraints(x), "constraints")
    assert_close(S.jacobian(x), O.jacobian(x), "jacobian")
    assert_close(S.hessian(x, lam, 0.7), O.hessian(x, lam, 0.7), "hessian")
