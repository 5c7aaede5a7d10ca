"""Initial-guess containers.  Only what the callback path and the solver
adapters touch: a view over one phase vector and the two guess builders
(``pockit/base/variablebase.py:92-131, 393-470``).  Interpolation / mesh
adaptation are out of scope."""
from __future__ import annotations

import numpy as np

from .phase import BcType, Phase

__all__ = ["Variable", "constant_guess", "linear_guess"]


class _Slabs:
    def __init__(self, data, lo, hi):
        self._d, self._lo, self._hi = data, lo, hi

    def __len__(self):
        return len(self._lo)

    def __getitem__(self, i):
        return self._d[self._lo[i] : self._hi[i]]

    def __setitem__(self, i, v):
        self._d[self._lo[i] : self._hi[i]] = v


class Variable:
    def __init__(self, phase: Phase, data: np.ndarray):
        if len(data) != phase.L:
            raise ValueError("data must have the same length as the phase vector")
        self.phase = phase
        self.data = np.asarray(data, dtype=np.float64)
        n_x = phase.n_x
        self.x = _Slabs(self.data, phase.l_v[:n_x], phase.r_v[:n_x])
        self.u = _Slabs(self.data, phase.l_v[n_x:], phase.r_v[n_x:])
        col = phase.col
        self._t_x = col.t_m if col.scheme == "lgl" else np.concatenate([col.t_m, [1.0]])

        self._t_u = col.t_m

    # physical times of the state / control nodes (``variablebase.py:355-363``)
    t_x = property(lambda self: self._t_x * (self.t_f - self.t_0) + self.t_0)
    t_u = property(lambda self: self._t_u * (self.t_f - self.t_0) + self.t_0)
    t_0 = property(lambda self: self.data[-2], lambda self, v: self.data.__setitem__(-2, v))
    t_f = property(lambda self: self.data[-1], lambda self, v: self.data.__setitem__(-1, v))


def _times(v: Variable, phase: Phase):
    if phase.info_t_0.t == BcType.FIXED:
        v.t_0 = phase.t_0
    else:
        v.t_0 -= 0.5
    if phase.info_t_f.t == BcType.FIXED:
        v.t_f = phase.t_f
    else:
        v.t_f += 0.5


def constant_guess(phase: Phase, value: float = 1.0) -> Variable:
    if not phase.ok:
        raise ValueError("phase is not fully configured")
    v = Variable(phase, np.full(phase.L, float(value)))
    for i in range(phase.n_x):
        if phase.info_bc_0[i].t == BcType.FIXED:
            v.x[i][0] = phase.bc_0[i]
        if phase.info_bc_f[i].t == BcType.FIXED:
            v.x[i][-1] = phase.bc_f[i]
    _times(v, phase)
    return v


def linear_guess(phase: Phase, default: float = 1.0) -> Variable:
    if not phase.ok:
        raise ValueError("phase is not fully configured")
    v = Variable(phase, np.full(phase.L, float(default)))
    for i in range(phase.n_x):
        f0 = phase.info_bc_0[i].t == BcType.FIXED
        f1 = phase.info_bc_f[i].t == BcType.FIXED
        if f0 and f1:
            v.x[i] = v._t_x * (phase.bc_f[i] - phase.bc_0[i]) + phase.bc_0[i]
        elif f0:
            v.x[i] = phase.bc_0[i]
        elif f1:
            v.x[i] = phase.bc_f[i]
    _times(v, phase)
    return v
