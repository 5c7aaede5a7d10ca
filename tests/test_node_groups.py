"""CPU tier: expression groups in the per-node programs (DevicePlan(node_groups=k): up to k threads
per node, each evaluating the functions of its group with their own CSE) must give the reference's
values like the one-thread-per-node programs do (host-emulated)."""
import numpy as np
import pytest

from helpers import assert_close, build, golden_cases, load
from hostemu import HostEmu
from pockit_b200 import plan as P


@pytest.mark.parametrize("groups", [2, 4, 16])
@pytest.mark.parametrize("case", ["general_lgl", "general_lgr", "robot_arm_lgr_6x20", "rocket_lgl_4x5", "humanoid_lgl_4x5",
                                  "quadrotor_lgl_14x6", "quadrotor_lgr_5x3", "no_control_lgr_3x3", "static_only_lgl", "tiny_lgl_1x3"])
def test_grouped_programs_match_reference(case, groups):
    S, g = build(case), load(case)
    E = HostEmu(S, node_groups=groups)
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    assert_close(E.run(P.OBJ, x)[0], g["objective"], "objective")
    assert_close(E.run(P.CONS, x), g["constraints"], "constraints")
    assert_close(E.run(P.GRAD, x), g["gradient"], "gradient")
    assert_close(E.run(P.JAC, x), g["jacobian"], "jacobian")
    assert_close(E.run(P.HESS, x, lam, sigma), g["hessian"], "hessian")
    assert_close(E.run(P.SET, x, lam, sigma)[-len(g["hessian"]):], g["hessian"], "hessian (set pipeline)")


def test_groups_split_the_functions():
    S = build("robot_arm_lgr_6x20")
    dp = P.DevicePlan(S.lowering, node_groups=4)
    mp = dp.mode(P.HESS)
    dp.source(P.HESS)
    groups = dp._function_groups(mp, 0)
    assert 2 <= len(groups) <= 4 and mp.node_threads[0] == len(groups)
    members = [fn for g in groups for fn in g]
    assert len(members) == len(set(members))  # every function in exactly one group
    assert dp.finalize(P.HESS)["node_threads"] == [len(groups)]
    one = P.DevicePlan(S.lowering)
    one.source(P.HESS)
    assert one.finalize(P.HESS)["node_threads"] == [1]


def test_batched_grouped_programs():
    import pockit_b200.lobatto as lob
    from pockit_b200 import problems

    S = problems.quadrotor(lob, mesh=3, num_point=4, fastmath=False)
    B = 3
    rng = np.random.default_rng(2)
    x0, lam0, _ = problems.evaluation_point(S)
    X = x0[None, :] + 1e-2 * rng.normal(size=(B, len(x0)))
    LAM = lam0[None, :] + 0.1 * rng.normal(size=(B, len(lam0)))
    sig = np.array([0.5, 1.0, 1.5])
    a, b = HostEmu(S, batch=B), HostEmu(S, batch=B, node_groups=3)
    for m, args in ((P.JAC, (X,)), (P.HESS, (X, LAM, sig)), (P.CONS, (X,)), (P.GRAD, (X,))):
        assert_close(b.run(m, *args), a.run(m, *args), P.MODES[m])
