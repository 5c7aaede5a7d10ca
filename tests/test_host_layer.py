"""Host layer against the known answers the reference's own tests hold:
sparse derivative index order (tests/test_base/test_fastfunc.py:11-71), LGL / LGR tables
(tests/test_labatto/test_discretization_lobatto.py:5-88, tests/test_radau/test_discretization_radau.py:5-60),
bounds (tests/test_labatto/test_bound_lobatto.py, tests/test_radau/test_bound_radau.py) and the
validation behaviour of the modelling API (tests/test_base/test_system_base.py:83-118)."""
import numpy as np
import pytest
import sympy as sp

import pockit_b200.lobatto as lob
import pockit_b200.radau as rad
from pockit_b200.discretization import Collocation, _unit_integration_block, gauss_lobatto, gauss_radau
from pockit_b200.symfunc import SymFunc


def _eval(exprs, syms, vals):
    return np.array([[float(e.subs(dict(zip(syms, v)))) for v in zip(*vals)] for e in exprs])


def test_symfunc_constant_has_no_derivatives():
    for args in ([], [sp.Symbol("x")]):
        f = SymFunc(1, args)
        assert f.expr == 1 and f.n_G == 0 and f.n_H == 0
        assert len(f.G_index) == len(f.H_index_row) == len(f.H_index_col) == 0


def test_symfunc_index_order_and_values():
    x, y = sp.symbols("x, y")
    t = np.arange(10, dtype=np.float64)
    vx, vy = np.sin(t), t * 2
    f = SymFunc(x + y**2, [x, y])
    assert f.G_index.tolist() == [0, 1] and f.n_H == 1  # d2/dy2 = 2
    f = SymFunc(x * y + y**3, [x, y])
    assert f.G_index.tolist() == [0, 1]
    assert f.H_index_row.tolist() == [1, 1] and f.H_index_col.tolist() == [0, 1]  # (r, c) in order
    assert np.allclose(_eval(f.G_expr, [x, y], [vx, vy]), np.vstack([vy, 3 * vy**2 + vx]))
    assert np.allclose(_eval(f.H_expr, [x, y], [vx, vy]), np.vstack([np.ones(10), 6 * vy]))
    f = SymFunc(x**2 * y + y**3, [x, y])
    assert f.H_index_row.tolist() == [0, 1, 1] and f.H_index_col.tolist() == [0, 0, 1]
    assert np.allclose(_eval(f.H_expr, [x, y], [vx, vy]), np.vstack([2 * vy, 2 * vx, 6 * vy]))


def test_gauss_lobatto_known_values():
    assert np.allclose(gauss_lobatto(1)[0], [0.0]) and np.allclose(gauss_lobatto(1)[1], [2.0])
    assert np.allclose(gauss_lobatto(2)[0], [-1.0, 1.0]) and np.allclose(gauss_lobatto(2)[1], [1.0, 1.0])
    x, w = gauss_lobatto(4)
    assert np.allclose(x, [-1.0, -1 / np.sqrt(5), 1 / np.sqrt(5), 1]) and np.allclose(w, [1 / 6, 5 / 6, 5 / 6, 1 / 6])
    x, w = gauss_lobatto(5)
    assert np.allclose(x, [-1.0, -np.sqrt(3 / 7), 0, np.sqrt(3 / 7), 1])
    assert np.allclose(w, [1 / 10, 49 / 90, 32 / 45, 49 / 90, 1 / 10])
    x, w = gauss_lobatto(10)
    assert np.allclose(x[1:5], [-0.9195339081664588138289, -0.7387738651055050750031, -0.4779249498104444956612,
                                -0.1652789576663870246262])
    assert np.allclose(w[:3], [0.02222222222222222222222, 0.1333059908510701111262, 0.2248893420631264521195])


def test_gauss_radau_known_values():
    assert np.allclose(gauss_radau(1)[0], [-1.0]) and np.allclose(gauss_radau(1)[1], [2.0])
    x, w = gauss_radau(3)
    assert np.allclose(x, [-1.0, -0.289898, 0.689898], atol=1e-6)
    assert np.allclose(w, [0.222222, 1.02497, 0.752806], atol=1e-5)
    x, w = gauss_radau(5)
    assert np.allclose(x, [-1.0, -0.72048, -0.167181, 0.446314, 0.885792], atol=1e-6)
    assert np.allclose(w, [0.08, 0.446208, 0.623653, 0.562712, 0.287427], atol=1e-6)


@pytest.mark.parametrize("n,f,F", [(10, lambda x: 2 * x, lambda x: x**2), (10, np.cos, np.sin),
                                    (20, lambda x: 10 * np.exp(x), lambda x: 10 * np.exp(x))])
def test_integration_blocks_integrate_from_the_right_end(n, f, F):
    x, _ = gauss_lobatto(n)
    assert np.allclose(_unit_integration_block("lgl", n) @ f(x), (F(x) - F(1.0))[:-1])
    x, _ = gauss_radau(n)
    assert np.allclose(_unit_integration_block("lgr", n) @ f(x), F(x) - F(1.0))


def test_layout_tables_for_a_two_interval_mesh():
    mesh, npt = np.array([0.0, 0.1, 1.0]), np.array([2, 3], dtype=np.int32)
    c = Collocation("lgl", mesh, npt, 2, 1)
    assert c.L_m == 4 and c.L_x == 4 and c.L == 14 and c.n_rows == 3
    assert c.l_v.tolist() == [0, 4, 8] and c.r_v.tolist() == [4, 8, 12] and c.l_d.tolist() == [0, 3]
    assert np.allclose(c.t_m, [0.0, 0.1, 0.55, 1.0]) and np.isclose(c.w_m.sum(), 1.0)
    y = np.array([3.0, 5.0, -2.0, 7.0])
    assert np.allclose(c.T.dot(y), [3.0 - 5.0, 5.0 - 7.0, -2.0 - 7.0])
    c = Collocation("lgr", mesh, npt, 2, 1)
    assert c.L_m == 5 and c.L_x == 6 and c.L == 19 and c.n_rows == 5
    assert c.l_v.tolist() == [0, 6, 12] and c.r_v.tolist() == [6, 12, 17] and c.l_d.tolist() == [0, 5]
    y = np.arange(6.0) ** 2
    assert np.allclose(c.T.dot(y), [y[0] - y[2], y[1] - y[2], y[2] - y[5], y[3] - y[5], y[4] - y[5]])
    assert len(c.I.f) == 2 and len(c.I.b) == 0 and len(c.T.b) == 3  # front column of I, terminal column of T


@pytest.mark.parametrize("mod", [lob, rad])
def test_variable_and_constraint_bounds(mod):
    s = mod.System(4)
    p = s.new_phase(2, 2)
    p.set_dynamics([0, 0]).set_boundary_condition([0, 0], [s.s[0], 0], None, s.s[2]).set_discretization(
        [0, 0.2, 1], [3, 4]
    ).set_phase_constraint([p.x[0], p.u[1], p.t, p.s[3]], [2, 4, 6, 8], [3, np.inf, 7, 9])
    s.set_phase([p]).set_objective(0).set_system_constraint([s.s[1]], [0], [1])
    nx = p.col.L_x
    nu = p.col.L_m
    lb = [2] * nx + [-np.inf] * nx + [-np.inf] * nu + [4] * nu + [6] * 2 + [2, 0, 6, 8]
    ub = [3] * nx + [np.inf] * nx + [np.inf] * nu + [np.inf] * nu + [7] * 2 + [3, 1, 7, 9]
    assert np.allclose(lb, s.v_lb) and np.allclose(ub, s.v_ub)
    # bounds on a FUNC boundary value / FUNC time that is a bare static parameter end up as bounds
    # of that parameter, not as system constraints (systembase.py:291-340)
    assert len(s.F_c) == 0 and len(s.c_lb) == 2 * p.col.n_rows


@pytest.mark.parametrize("mod,least", [(rad, 1), (lob, 2)])
def test_discretization_validation_is_atomic(mod, least):
    system = mod.System(0)
    phase = system.new_phase(1, 0)
    phase.set_dynamics([0]).set_boundary_condition([0], [0], 0, 1)
    phase.set_discretization(1, max(least, 3))
    before = (phase._mesh.copy(), phase._num_point.copy(), phase.col)
    for mesh, npt in [(0, 3), ([0], [3]), ([0, 0], [3]), ([1, 0], [3]), ([0, np.inf], [3]), ([0, 0.5, 1], [3]),
                      ([0, 1], [least - 1]), ([0, 1], [2.5])]:
        with pytest.raises(ValueError):
            phase.set_discretization(mesh, npt)
        assert phase.ok and np.array_equal(phase._mesh, before[0]) and np.array_equal(phase._num_point, before[1])
        assert phase.col is before[2]


def test_model_api_errors():
    s = lob.System(1)
    with pytest.raises(ValueError):
        lob.System(1.5)
    with pytest.raises(ValueError):
        s.new_phase(["t"], 1)
    p = s.new_phase(2, 1)
    with pytest.raises(ValueError):
        p.set_dynamics([0])
    with pytest.raises(ValueError):
        p.set_phase_constraint([p.x[0]], [0, 1], [1])
    with pytest.raises(ValueError):
        p.set_phase_constraint([p.u[0]], [0], [np.inf], True)
    with pytest.raises(ValueError):
        p.set_boundary_condition([0], [0], 0, 1)
    with pytest.raises(ValueError):
        p.set_boundary_condition(["a", 0], [0, 0], 0, 1)
    with pytest.raises(ValueError):
        s.set_phase([p])  # phase not fully configured
    with pytest.raises(ValueError):
        s.set_system_constraint([s.s[0]], [0, 1], [1])


def test_callbacks_need_the_cuda_engine(monkeypatch):
    """No CPU fallback: without the library (or without a device) the callbacks raise."""
    import pockit_b200.engine as eng
    from pockit_b200 import problems

    S = problems.lqr(lob, 3, 3)
    monkeypatch.setattr(eng, "_LIB", None)
    monkeypatch.setenv("POCKIT_B200_LIB", "/nonexistent/libpockit_b200.so")
    with pytest.raises(RuntimeError):
        S.objective(np.zeros(S.L))


def test_generated_source_does_not_depend_on_the_mesh():
    """Sizes and offsets reach the per-node programs through a table, so re-meshing a model yields the
    same CUDA source (the engine's cubin cache then compiles nothing)."""
    import pockit_b200.radau as rad
    from pockit_b200 import plan as P
    from pockit_b200 import problems

    def sources(S):
        dp = P.DevicePlan(S.lowering)
        for m in range(6):
            dp.mode(m)
        return [dp.finalize(m)["source"] for m in range(6)]

    base = sources(problems.robot_arm(rad, 6, 20))
    assert sources(problems.robot_arm(rad, 9, 7)) == base
    assert sources(problems.robot_arm(rad, [0, 0.3, 0.5, 1.0], [4, 6, 5])) == base
