"""CPU tier: the fused set pipeline (plan.SET -- one per-node program, one reduction, one system
program for all five callbacks) must produce, slice by slice, what the five separate callback
plans produce (host-emulated) and the reference's values.  One CSE over the union of the leaves may
regroup a product differently from the per-callback CSE, so the two agree to the parity tolerance
(1e-12 rel / 1e-14 abs; observed: last-bit differences), not necessarily bit for bit."""
import numpy as np
import pytest

from helpers import assert_close, build, golden_cases, load
from hostemu import HostEmu
from pockit_b200 import plan as P

CASES = sorted(golden_cases())


@pytest.mark.parametrize("case", CASES)
def test_set_pipeline_equals_the_five_plans(case):
    S, g = build(case), load(case)
    E = HostEmu(S)
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    out = E.run(P.SET, x, lam, sigma)
    sub = E.fin[P.SET]["sub_range"]
    assert [sub[m][0] for m in P.SET_ORDER] == list(np.cumsum([0] + [sub[m][1] for m in P.SET_ORDER[:-1]]))
    assert sum(c for _, c in sub.values()) == len(out) and not np.isnan(out).any()
    single = {P.OBJ: E.run(P.OBJ, x), P.GRAD: E.run(P.GRAD, x), P.CONS: E.run(P.CONS, x), P.JAC: E.run(P.JAC, x),
              P.HESS: E.run(P.HESS, x, lam, sigma)}
    for m in P.SET_ORDER:
        off, cnt = sub[m]
        assert_close(out[off : off + cnt], single[m], P.MODES[m])
    for m, name in ((P.OBJ, "objective"), (P.GRAD, "gradient"), (P.CONS, "constraints"), (P.JAC, "jacobian"), (P.HESS, "hessian")):
        off, cnt = sub[m]
        assert_close(out[off : off + cnt] if m != P.OBJ else out[off], g[name], name)
    # the shared per-node program is not larger than the five separate ones together
    assert len(E.fin[P.SET]["source"]) < sum(len(E.fin[m]["source"]) for m in range(5))


def test_set_pipeline_batched():
    import pockit_b200.lobatto as lob
    from pockit_b200 import problems

    S = problems.quadrotor(lob, mesh=3, num_point=4, fastmath=False)
    B = 3
    rng = np.random.default_rng(1)
    x0, lam0, _ = problems.evaluation_point(S)
    X = x0[None, :] + 1e-2 * rng.normal(size=(B, len(x0)))
    LAM = lam0[None, :] + 0.1 * rng.normal(size=(B, len(lam0)))
    sig = np.array([0.5, 1.0, 1.5])
    E = HostEmu(S, batch=B)
    out = E.run(P.SET, X, LAM, sig)
    sub = E.fin[P.SET]["sub_range"]
    for m, args in ((P.GRAD, (X,)), (P.CONS, (X,)), (P.JAC, (X,)), (P.HESS, (X, LAM, sig))):
        off, cnt = sub[m]
        assert_close(out[:, off : off + cnt], E.run(m, *args), P.MODES[m])
    assert_close(out[:, 0], E.run(P.OBJ, X).reshape(-1), "objective")


@pytest.mark.parametrize("case", ["general_lgl", "general_lgr", "rocket_lgl_4x5", "robot_arm_lgr_6x20", "quadrotor_lgl_14x6",
                                  "humanoid_lgl_4x5", "lqr_lgl_10x10", "static_only_lgl", "tiny_lgl_1x3"])
def test_small_set_pipeline_covers_the_three_small_callbacks(case):
    """The engine's default: the pipeline covers objective, gradient and constraints only (plan.SMALL_SET);
    the Jacobian and the Hessian keep their own plans."""
    S, g = build(case), load(case)
    E = HostEmu(S, set_subs=P.SMALL_SET)
    x = g["x"]
    out = E.run(P.SET, x)
    sub = E.fin[P.SET]["sub_range"]
    assert sorted(sub) == sorted(P.SMALL_SET) and sum(c for _, c in sub.values()) == len(out)
    assert sub[P.OBJ] == (0, 1) and sub[P.GRAD] == (1, S.L) and sub[P.CONS] == (1 + S.L, len(S.c_lb))
    assert_close(out[0], g["objective"], "objective")
    assert_close(out[1 : 1 + S.L], g["gradient"], "gradient")
    assert_close(out[1 + S.L :], g["constraints"], "constraints")
    # no Jacobian / Hessian leaves in the shared program
    assert len(E.fin[P.SET]["source"]) < len(E.fin[P.HESS]["source"]) + len(E.fin[P.GRAD]["source"]) + len(E.fin[P.CONS]["source"])
