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
        # hp-refined style meshes: runs of equal order -> several block pieces, mixed unit blocks
        ("robot_arm", "radau", dict(mesh=24, num_point=[3] * 5 + [6] * 7 + [4] * 3 + [9] * 6 + [5, 3, 3])),
        ("rocket", "lobatto", dict(mesh=20, num_point=[4] * 6 + [7] * 8 + [3, 5, 5, 5, 8, 8])),
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


def test_batched_instances_match_per_instance_oracle():
    """BASELINE configs[4] in small: quadrotor LGL 14x6, instances differing only in the
    FIXED initial state; every instance must match an oracle built for that instance."""
    import pockit_b200.lobatto as lob
    from oracle.pockit_oracle import OracleSystem
    from pockit_b200 import problems
    from pockit_b200.batched import BatchedSystem, fixed_index, fixed_table

    B = 6
    rng = np.random.default_rng(0)
    starts = rng.uniform(-0.2, 0.2, size=(B, 2))
    S = problems.quadrotor(lob, fastmath=False)
    fixed = fixed_table(S, B)
    fixed[:, fixed_index(S, 0, "x0", 0)] = starts[:, 0]
    fixed[:, fixed_index(S, 0, "x0", 1)] = starts[:, 1]
    x0, lam0, _ = problems.evaluation_point(S)
    X = x0[None, :] + 1e-2 * rng.normal(size=(B, len(x0)))
    LAM = lam0[None, :] + 0.1 * rng.normal(size=(B, len(lam0)))
    sig = rng.uniform(0.5, 1.5, B)
    bs = BatchedSystem(S, fixed)
    obj, grad, cons = bs.objective(X), bs.gradient(X), bs.constraints(X)
    jac, hess = bs.jacobian(X), bs.hessian(X, LAM, sig)
    assert jac.shape == (B, len(S.jacobianstructure()[0]))
    for b in range(B):
        O = OracleSystem(problems.quadrotor(lob, start=tuple(starts[b]), fastmath=False))
        assert_close(obj[b], O.objective(X[b]), f"objective[{b}]")
        assert_close(grad[b], O.gradient(X[b]), f"gradient[{b}]")
        assert_close(cons[b], O.constraints(X[b]), f"constraints[{b}]")
        assert_close(jac[b], O.jacobian(X[b]), f"jacobian[{b}]")
        assert_close(hess[b], O.hessian(X[b], LAM[b], sig[b]), f"hessian[{b}]")
    bs.close()


def test_size_independent_properties_at_full_size():
    """BASELINE configs[1] at full size (40 000 nodes): properties that need no oracle run.
    Hessian values are linear in (lambda, sigma); the Jacobian is the derivative of the
    constraints (directional finite difference); callbacks are idempotent and leave x alone."""
    import pockit_b200.radau as rad
    from pockit_b200 import problems

    S = problems.robot_arm(rad, 2000, 20)
    x, lam, _ = problems.evaluation_point(S)
    assert S.L == 360008 and len(S.c_lb) == 240000
    jr, jc = S.jacobianstructure()
    hr, hc = S.hessianstructure()
    assert len(jr) == 12479754 and len(hr) == 12799740
    assert np.all(hr >= hc)  # lower triangle
    x_in = x.copy()
    j1, j2 = S.jacobian(x), S.jacobian(x)
    assert np.array_equal(j1, j2) and np.array_equal(x, x_in)
    rng = np.random.default_rng(3)
    lam2 = rng.normal(size=len(lam))
    h_a = S.hessian(x, lam, 1.0)
    h_b = S.hessian(x, lam2, 0.0)
    h_ab = S.hessian(x, 2.0 * lam + lam2, 2.0)
    np.testing.assert_allclose(h_ab, 2.0 * h_a + h_b, rtol=1e-10, atol=1e-10)
    d = rng.normal(size=S.L)
    eps = 1e-6
    fd = (S.constraints(x + eps * d) - S.constraints(x - eps * d)) / (2 * eps)
    jd = np.zeros(len(lam))
    np.add.at(jd, jr, j1 * d[jc])
    np.testing.assert_allclose(jd, fd, rtol=1e-5, atol=1e-6)
    g = S.gradient(x)
    fdo = (S.objective(x + eps * d) - S.objective(x - eps * d)) / (2 * eps)
    np.testing.assert_allclose(g @ d, fdo, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("pipeline", ["0", "1", "small"])
def test_evaluation_set_graph_matches_single_callbacks(pipeline, monkeypatch):
    """pk_run_set (all callbacks at one x as one graph) against the five separate host-to-host
    callbacks.  POCKIT_B200_SET=0: one stream per callback, the same kernels -> bit-identical.
    POCKIT_B200_SET=small (the default): objective / gradient / constraints as one pipeline (its single
    CSE may regroup a product -> within the parity tolerance), Jacobian / Hessian bit-identical.
    POCKIT_B200_SET=1: one pipeline for all five."""
    import pockit_b200.lobatto as lob
    from pockit_b200 import plan as P
    from pockit_b200 import problems

    monkeypatch.setenv("POCKIT_B200_SET", pipeline)
    S = problems.rocket(lob, mesh=30, num_point=8)
    x, lam, sigma = problems.evaluation_point(S, seed=5)
    single = {
        P.OBJ: np.atleast_1d(S.objective(x)), P.GRAD: S.gradient(x), P.CONS: S.constraints(x),
        P.JAC: S.jacobian(x), P.HESS: S.hessian(x, lam, sigma),
    }
    eng = S.engine
    assert eng.has_set == (pipeline != "0")
    covered = {"0": (), "1": P.SET_ORDER, "small": P.SMALL_SET}[pipeline]  # "small" is the engine's default
    modes = [P.OBJ, P.GRAD, P.CONS, P.JAC, P.HESS]
    eng.upload(x, lam, sigma)
    for _ in range(3):  # replays of the captured graph
        eng.run_set(modes)
    eng.sync()
    for m in modes:
        got = np.atleast_1d(eng.download(m))
        if m not in covered:
            assert np.array_equal(got, single[m]), P.MODES[m]
        else:
            assert_close(got, single[m], P.MODES[m])
    # a single callback afterwards takes over its output again
    assert np.array_equal(S.jacobian(x + 1e-3), np.atleast_1d(eng.download(P.JAC)))


@pytest.mark.parametrize("name", ["general_lgl", "general_lgr", "rocket_lgl_4x5", "robot_arm_lgr_6x20", "quadrotor_lgl_14x6",
                                  "humanoid_lgl_4x5", "lqr_lgl_10x10", "static_only_lgl", "no_control_lgr_3x3", "tiny_lgl_1x3"])
def test_set_pipeline_matches_reference_golden(name, monkeypatch):
    """System.evaluate through the opt-in fused set pipeline (POCKIT_B200_SET=1) against the real
    reference's values."""
    monkeypatch.setenv("POCKIT_B200_SET", "1")
    S, g = build(name), load(name)
    x, lam, sigma = g["x"], g["lam"], float(g["sigma"])
    r = S.evaluate(x.copy(), lam, sigma)
    assert S.engine.has_set
    for k in ("objective", "gradient", "constraints", "jacobian", "hessian"):
        assert_close(r[k], g[k], k)


def test_fastmath_model_compiles_and_stays_close():
    """fastmath=True (the reference passes it to Numba; here it maps to --fmad=true): results may
    differ in the last bits but must stay within a few ulp of the strict evaluation."""
    import pockit_b200.lobatto as lob
    from pockit_b200 import problems

    strict = problems.quadrotor(lob, fastmath=False)
    fast = problems.quadrotor(lob, fastmath=True)
    x, lam, sigma = problems.evaluation_point(strict)
    np.testing.assert_allclose(fast.jacobian(x), strict.jacobian(x), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(fast.hessian(x, lam, sigma), strict.hessian(x, lam, sigma), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(fast.constraints(x), strict.constraints(x), rtol=1e-13, atol=1e-13)
