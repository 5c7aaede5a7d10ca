"""Guess <-> optimisation vector conversion shared by the adapters
(behaviour of ``pockit/optimizer/_common.py:9-63``: same checks, same ``ValueError`` texts)."""
from __future__ import annotations

import numpy as np

from ..guess import Variable


def pack_guess(system, guess, optimizer_options):
    if not system.ok:
        raise ValueError("system is not fully configured")
    single = isinstance(guess, Variable)
    parts = [guess] if single else list(guess)
    if not system.n_s and len(parts) != system.n_p:
        raise ValueError("len(guess) must be equal to the number of phases")
    if system.n_s and len(parts) != system.n_p + 1:
        raise ValueError("len(guess) must be equal to the number of phases + 1 (for static variables)")
    x0 = np.zeros(system.L)
    for i in range(system.n_p):
        x0[system.l_p[i] : system.r_p[i]] = parts[i].data
    if system.n_s:
        x0[system.l_s : system.r_s] = np.array(list(parts[-1]), dtype=np.float64)
    return x0, single, dict(optimizer_options or {})


def unpack_solution(system, x, single):
    """Phase vectors with the FIXED / FUNC boundary values substituted (the solver never sees them
    move), wrapped as ``Variable`` objects; static parameters last."""
    x = np.array(x, dtype=np.float64)
    s = x[system.l_s : system.r_s]
    out = []
    for i, p in enumerate(system.p):
        xp = x[system.l_p[i] : system.r_p[i]]
        for j in range(p.n_x):
            xp[p.l_v[j]] = p._value_boundary_condition(p.info_bc_0[j], xp[p.l_v[j]], s)
            xp[p.r_v[j] - 1] = p._value_boundary_condition(p.info_bc_f[j], xp[p.r_v[j] - 1], s)
        xp[-2] = p._value_boundary_condition(p.info_t_0, xp[-2], s)
        xp[-1] = p._value_boundary_condition(p.info_t_f, xp[-1], s)
        out.append(Variable(p, xp))
    if system.n_s:
        out.append(s)
    return out[0] if single else out
