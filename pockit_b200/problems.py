"""Workload generators: the BASELINE.json configurations re-stated with a
parametrised scheme / mesh.  Every builder takes the transcription
module (``pockit_b200.lobatto``, ``pockit_b200.radau`` -- or the reference's
``pockit.lobatto`` / ``pockit.radau`` when generating golden vectors) so the very
same model text drives both sides of a parity test.

Model sources (equations and constants only): ``examples/linear_quadratic_regulator.py:47-62``,
``examples/robot_arm.py:62-138``, ``examples/humanoid_whole_body_control.py:157-275``,
``examples/multiphase_two_stage_rocket.py:48-155``, ``examples/planar_quadrotor.py:57-140``,
and the general derivative-test system ``tests/test_radau/test_derivative_radau.py:11-41``.
"""
from __future__ import annotations

import numpy as np
import sympy as sp

__all__ = [
    "lqr", "robot_arm", "humanoid", "rocket", "quadrotor", "general", "static_only", "no_control", "tiny",
    "evaluation_point", "BUILDERS",
]


def lqr(mod, mesh=10, num_point=10):
    S = mod.System(["x_final"])
    (xf,) = S.s
    p = S.new_phase(["x"], ["u"])
    (x,), (u,) = p.x, p.u
    p.set_dynamics([-1.0 * x + 1.0 * u])
    p.set_integral([1.0 * x**2 + 0.1 * u**2])
    p.set_boundary_condition([1.0], [xf], 0.0, 1.0)
    p.set_discretization(mesh, num_point)
    S.set_phase([p])
    S.set_objective(p.I[0] + 1.0 * xf**2 / 2.0)
    return S


def robot_arm(mod, mesh=2000, num_point=20):
    arm = 5.0
    S = mod.System(0)
    p = S.new_phase(
        ["pivot_position", "pivot_speed", "azimuth", "azimuth_rate", "polar_angle", "polar_angle_rate"],
        ["pivot_force", "azimuth_torque", "polar_torque"],
    )
    rho, rho_dot, _, az_dot, phi, phi_dot = p.x
    f_rho, tq_az, tq_phi = p.u
    i_phi = ((arm - rho) ** 3 + rho**3) / 3.0
    i_az = i_phi * sp.sin(phi) ** 2
    p.set_dynamics([rho_dot, f_rho / arm, az_dot, tq_az / i_az, phi_dot, tq_phi / i_phi])
    p.set_integral([1.0])
    margin = np.deg2rad(10.0)
    p.set_phase_constraint(
        [rho, phi, f_rho, tq_az, tq_phi],
        [0.0, margin, -1.0, -1.0, -1.0],
        [arm, np.pi - margin, 1.0, 1.0, 1.0],
        [False, False, True, True, True],
    )
    p.set_boundary_condition(
        [4.5, 0.0, 0.0, 0.0, np.pi / 4.0, 0.0],
        [4.5, 0.0, 2.0 * np.pi / 3.0, 0.0, np.pi / 4.0, 0.0],
        0.0,
        None,
    )
    p.set_discretization(mesh, num_point)
    S.set_phase([p])
    S.set_objective(p.I[0])
    return S


def _hands(q):
    torso_len, upper, fore = 0.60, 0.38, 0.30
    _, rs, re, ls, le = q
    sh = sp.Matrix([0.0, torso_len])
    right = sh + sp.Matrix(
        [upper * sp.cos(rs) + fore * sp.cos(rs + re), upper * sp.sin(rs) + fore * sp.sin(rs + re)]
    )
    left = sh + sp.Matrix(
        [-upper * sp.cos(ls) - fore * sp.cos(ls + le), upper * sp.sin(ls) + fore * sp.sin(ls + le)]
    )
    return right, left


def humanoid(mod, mesh=1000, num_point=10):
    horizon, kp, kd = 2.5, 36.0, 12.0
    q0 = np.array([0.25, -0.45, 0.95, 0.60, -1.10])
    q_left = np.array([0.0, -0.70, 1.10, -0.45, 1.00])
    disp = np.array([0.04, 0.08])

    def numeric_hand(q, which):
        r, l = _hands(sp.Matrix(q))
        return np.array((r if which == "r" else l).evalf(), dtype=float).ravel()

    right0 = numeric_hand(q0, "r")
    left_target = numeric_hand(q_left, "l")

    S = mod.System(0)
    p = S.new_phase(
        ["torso_angle", "right_shoulder_angle", "right_elbow_angle", "left_shoulder_angle",
         "left_elbow_angle", "torso_rate", "right_shoulder_rate", "right_elbow_rate",
         "left_shoulder_rate", "left_elbow_rate"],
        ["null_torso_acceleration", "null_right_shoulder_acceleration", "null_right_elbow_acceleration",
         "null_left_shoulder_acceleration", "null_left_elbow_acceleration"],
    )
    q = sp.Matrix(p.x[:5])
    qd = sp.Matrix(p.x[5:])
    z = sp.Matrix(p.u)
    right, left = _hands(q)
    j_r = right.jacobian(q)
    j_l = left.jacobian(q)
    j_r_dot = sp.zeros(2, 5)
    for k in range(5):
        j_r_dot += j_r.diff(q[k]) * qd[k]
    a = j_r[:, 1:3]
    det = a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]
    a_inv = sp.Matrix([[a[1, 1], -a[0, 1]], [-a[1, 0], a[0, 0]]]) / det
    pinv = sp.zeros(5, 2)
    pinv[1:3, :] = a_inv
    null_proj = sp.diag(1.0, 0.0, 0.0, 1.0, 1.0)
    tau = p.t / horizon
    prog = 10.0 * tau**3 - 15.0 * tau**4 + 6.0 * tau**5
    prog_d = (30.0 * tau**2 - 60.0 * tau**3 + 30.0 * tau**4) / horizon
    prog_dd = (60.0 * tau - 180.0 * tau**2 + 120.0 * tau**3) / horizon**2
    want_pos = sp.Matrix(right0) + sp.Matrix(disp) * prog
    want_vel = sp.Matrix(disp) * prog_d
    want_acc = sp.Matrix(disp) * prog_dd
    a_ref = want_acc + kp * (want_pos - right) + kd * (want_vel - j_r * qd)
    qdd = pinv * (a_ref - j_r_dot * qd) + null_proj * z
    p.set_dynamics([*qd, *qdd])
    err = left - sp.Matrix(left_target)
    v_l = j_l * qd
    p.set_integral(
        [120.0 * err.dot(err) + 3.0 * v_l.dot(v_l) + 20.0 * q[0] ** 2 + 0.03 * qd.dot(qd) + 0.01 * z.dot(z)]
    )
    p.set_phase_constraint(
        [*q, *qd, *z],
        [-0.55, -1.8, 0.35, -2.2, -2.2, *([-3.0] * 5), *([-10.0] * 5)],
        [0.55, 1.2, 1.8, 2.2, 2.2, *([3.0] * 5), *([10.0] * 5)],
    )
    p.set_boundary_condition([*q0, *np.zeros(5)], [None] * 10, 0.0, horizon)
    p.set_discretization(mesh, num_point)
    S.set_phase([p])
    S.set_objective(p.I[0])
    return S


def rocket(mod, mesh=5556, num_point=10):
    m0, prop1, drop, prop2, h_target = 1.0, 0.06, 0.20, 0.12, 2.0
    burnout1 = m0 - prop1
    m2_0 = burnout1 - drop
    m2_dry = m2_0 - prop2
    S = mod.System(
        ["h_separation", "v_separation", "m_before_drop", "m_after_drop", "t_separation", "m_final", "t_final"]
    )
    h_s, v_s, m_b, m_a, t_s, m_f, t_f = S.s

    def stage(tag, thrust, flow, mass_bounds, bc0, bcf, t0, tf):
        p = S.new_phase([f"altitude_{tag}", f"velocity_{tag}", f"mass_{tag}"], [f"throttle_{tag}"])
        h, v, m = p.x
        (thr,) = p.u
        p.set_dynamics([v, thrust * thr / m - 1.0, -flow * thr])
        p.set_integral([thr**2])
        p.set_phase_constraint(
            [thr, h, v, m],
            [0.0, 0.0, 0.0, mass_bounds[0]],
            [1.0, h_target, 3.0, mass_bounds[1]],
            [True, False, False, False],
        )
        p.set_boundary_condition(bc0, bcf, t0, tf)
        p.set_discretization(mesh, num_point)
        return p

    p1 = stage(1, 2.40, 0.080, (burnout1, m0), [0.0, 0.0, m0], [h_s, v_s, m_b], 0.0, t_s)
    p2 = stage(2, 1.60, 0.045, (m2_dry, m2_0), [h_s, v_s, m_a], [h_target, 0.0, m_f], t_s, t_f)
    S.set_phase([p1, p2])
    S.set_objective(t_f + 0.04 * (p1.I[0] + p2.I[0]))
    S.set_system_constraint(
        [h_s, v_s, m_b, m_a, m_a - m_b, t_s, m_f, t_f - t_s, t_f],
        [0.10, 0.05, burnout1, m2_dry, -drop, 0.20, m2_dry, 0.40, 1.00],
        [1.80, 2.50, burnout1, m2_0, -drop, 3.00, m2_0, 5.00, 7.00],
    )
    return S


def quadrotor(mod, mesh=14, num_point=6, start=(0.0, 0.0), fastmath=True):
    mass, inertia, grav, horizon = 1.20, 0.025, 9.81, 5.0
    max_torque = 0.25
    S = mod.System(0, fastmath=fastmath)
    p = S.new_phase(["x", "z", "velocity_x", "velocity_z", "pitch", "pitch_rate"], ["thrust", "torque"])
    x, z, vx, vz, th, th_d = p.x
    thrust, torque = p.u
    p.set_dynamics(
        [vx, vz, -thrust * sp.sin(th) / mass, thrust * sp.cos(th) / mass - grav, th_d, torque / inertia]
    )
    hover = mass * grav
    p.set_integral(
        [0.025 * ((thrust - hover) / hover) ** 2 + 0.012 * (torque / max_torque) ** 2
         + 0.002 * th_d**2 + 0.004 * z**2]
    )
    d2 = (x - 2.5) ** 2 + (z - 0.80) ** 2
    nt = p.t / horizon
    guard = 16.0 * 0.03 * nt**2 * (1.0 - nt) ** 2
    radius = 0.80 + 0.12 + 0.004
    p.set_phase_constraint(
        [x, z, th, th_d, thrust, torque, d2, z - guard],
        [-0.20, 0.0, -np.deg2rad(65.0), -3.0, 0.0, -max_torque, radius**2, 0.0],
        [5.20, 3.0, np.deg2rad(65.0), 3.0, 2.2 * mass * grav, max_torque, np.inf, np.inf],
    )
    p.set_boundary_condition(
        [float(start[0]), float(start[1]), 0.0, 0.0, 0.0, 0.0], [5.0, 0.0, 0.0, 0.0, 0.0, 0.0], 0.0, horizon
    )
    p.set_discretization(mesh, num_point)
    S.set_phase([p])
    S.set_objective(p.I[0])
    return S


def general(mod, mesh=(0, 0.2, 1), num_point=(3, 4), linear_objective=False):
    """The reference's most general small system: 2 static parameters, FIXED / FUNC
    boundary values, free ``t_0``, FUNC ``t_f``, two integrals, two path
    constraints, non-uniform mesh.  With ``linear_objective`` the objective and
    system constraints stay first-order in the integrals."""
    S = mod.System(2)
    p = S.new_phase(1, 1)
    s0, s1 = S.s
    x, u, t = p.x[0], p.u[0], p.t
    p.set_dynamics([x * sp.cos(s0) / u + t**2])
    p.set_boundary_condition([0], [sp.cos(s0 * 0.1)], None, 3 * sp.sin(s1))
    p.set_integral(
        [
            sp.cos(x) * u + 2 * x * sp.cos(s0) + 3 * sp.cos(x) * t + 4 * u * sp.cos(s0)
            + 5 * sp.cos(u) * t + 6 * s1 * sp.cos(t),
            6 * sp.cos(x) * u + 5 * x * sp.cos(s0) + 4 * sp.cos(x) * t + 3 * u * sp.cos(s0)
            + 2 * sp.cos(u) * t + s1 * sp.cos(t),
        ]
    )
    p.set_phase_constraint([t - x * u * s0 * s1, x], [0, 0], [0, 1])
    p.set_discretization(list(mesh) if not isinstance(mesh, int) else mesh,
                         list(num_point) if not isinstance(num_point, int) else num_point)
    S.set_phase([p])
    if linear_objective:
        S.set_objective(p.I[0] * sp.cos(s0) + 2 * p.I[1] + s0**2 * s1)
        S.set_system_constraint([(s0 + 1) ** 2, s1 / 2 + p.I[0]], [0, 0], [0, 0])
    else:
        S.set_objective((p.I[0] + p.I[1] + s0) ** 2)
        S.set_system_constraint([(s0 + 1) ** 2, s1 / 2 * p.I[0]], [0, 0], [0, 0])
    return S


def static_only(mod):
    """No phase at all: objective and system constraints of static parameters only
    (``tests/test_base/test_system_base.py:11-19`` in the reference)."""
    S = mod.System(2)
    S.set_objective(S.s[0] ** 2 * sp.cos(S.s[1]) + S.s[1])
    S.set_system_constraint([S.s[0] * S.s[1], S.s[0]], [0, -1], [1, 1])
    return S


def no_control(mod, mesh=3, num_point=3):
    """A phase without controls, FUNC terminal value, explicit time dependence."""
    S = mod.System(1)
    p = S.new_phase(2, 0)
    p.set_dynamics([p.x[1], -p.x[0] * S.s[0] + p.t])
    p.set_boundary_condition([1.0, None], [None, S.s[0] ** 2], 0.0, 2.0)
    p.set_integral([p.x[0] ** 2])
    p.set_discretization(mesh, num_point)
    S.set_phase([p])
    S.set_objective(p.I[0] + S.s[0])
    return S


def tiny(mod, mesh=1, num_point=3):
    """Smallest meshes: one or two intervals, middle node sets of length 0 or 1."""
    S = mod.System(0)
    p = S.new_phase(1, 1)
    p.set_dynamics([p.x[0] * p.u[0]])
    p.set_boundary_condition([1.0], [None], 0.0, None)
    p.set_integral([p.u[0] ** 2 + p.x[0]])
    p.set_discretization(mesh, num_point)
    S.set_phase([p])
    S.set_objective(p.I[0])
    return S


def check_system(mod):
    """The reference's known-answer system for the continuous error check
    (``tests/test_labatto/test_check_lobatto.py:7-36``, same for radau): ``x' = u`` on a two-interval
    mesh, with trajectories that are exact polynomials (must pass) and perturbed ones (must fail).
    Returns ``(system, [(value, expected), ...])`` with ``value = [Variable, statics]``."""
    from .guess import constant_guess

    S = mod.System(1)
    p = S.new_phase(1, 1)
    p.set_dynamics([p.u[0]])
    p.set_boundary_condition([None], [None], None, None)
    p.set_phase_constraint([p.u[0] + S.s[0]], [0.0], [2.0], [True])
    p.set_discretization([0, 0.1, 1], [3, 4])
    S.set_phase([p])
    S.set_objective(S.s[0])

    def variable(x_of_t, u_of_t, bump=0.0):
        v = constant_guess(p, 1.0)
        v.x[0] = x_of_t(v.t_x)
        v.u[0] = u_of_t(v.t_u)
        v.u[0][0] += bump
        return [v, [0.0]]

    one = lambda t: np.ones_like(t)  # noqa: E731
    cases = [
        (variable(lambda t: t, one), True),
        (variable(lambda t: t**2, lambda t: 2 * t), True),
        (variable(lambda t: t**2, lambda t: 2 * t, bump=0.01), False),
        (variable(lambda t: t**2, lambda t: 1.99 * t), False),
    ]
    return S, cases


BUILDERS = {
    "static_only": static_only, "no_control": no_control, "tiny": tiny,
    "lqr": lqr, "robot_arm": robot_arm, "humanoid": humanoid, "rocket": rocket,
    "quadrotor": quadrotor, "general": general,
}


def evaluation_point(S, seed: int = 1, jitter: float = 1e-2):
    """Synthetic ``(x, lam, sigma)``: a smooth, singularity-free guess plus
    ``jitter * N(0, 1)``.  Works on either implementation."""
    rng = np.random.default_rng(seed)
    L = int(S.L)
    x = np.zeros(L)
    for i, p in enumerate(S.p):
        seg = np.full(int(p.L), 0.5)
        n_x = p.n_x
        for j in range(n_x):
            lo, hi = int(p.l_v[j]), int(p.r_v[j])
            b0, bf = p.bc_0[j], p.bc_f[j]
            a = float(b0) if isinstance(b0, float) else (0.6 if b0 is None or not isinstance(b0, float) else b0)
            b = float(bf) if isinstance(bf, float) else a + 0.3
            seg[lo:hi] = np.linspace(a, b, hi - lo)
        seg[-2] = float(p.t_0) if isinstance(p.t_0, float) else 0.1 + i
        seg[-1] = float(p.t_f) if isinstance(p.t_f, float) else 2.0 + i
        x[int(S.l_p[i]) : int(S.r_p[i])] = seg
    if S.n_s:
        x[int(S.l_s) : int(S.r_s)] = 0.7 + 0.1 * np.arange(S.n_s)
    x = x + jitter * rng.normal(size=L)
    m = len(S.c_lb)
    lam = np.random.default_rng(seed + 1).normal(size=m)
    return x, lam, 1.0
