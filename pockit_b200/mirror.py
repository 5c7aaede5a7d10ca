"""Build the engine-backed twin of a live reference system.

A model script that already constructs ``pockit.lobatto.System`` / ``pockit.radau.System`` objects can
keep doing so and hand the finished system over::

    from pockit_b200.mirror import from_reference
    fast = from_reference(system)            # same NLP, callbacks on the B200
    nlp = cyipopt.Problem(..., problem_obj=fast, ...)

Everything the twin needs is what the reference keeps from its own setters (``phasebase.py:242-630``,
``systembase.py:148-255``): the SymPy expressions of dynamics / integrals / phase constraints, the raw
boundary values, mesh and orders, objective and system constraints.  Symbols are mapped by position,
so the two systems share layout, bounds and COO patterns bit for bit (``tests/test_mirror.py``).
Nothing of the reference package is imported here: the argument is only read.
"""
from __future__ import annotations

import importlib

import numpy as np
import sympy as sp

__all__ = ["from_reference"]


def _scheme_of(ref_system) -> str:
    mod = type(ref_system).__module__
    if ".lobatto" in mod:
        return "lobatto"
    if ".radau" in mod:
        return "radau"
    raise ValueError(f"cannot tell the transcription of {type(ref_system)!r}; pass scheme='lobatto' or 'radau'")


def from_reference(ref_system, scheme: str | None = None):
    """``pockit_b200`` System equivalent to the (fully configured) reference ``ref_system``."""
    if not getattr(ref_system, "ok", False):
        raise ValueError("system is not fully configured")
    mod = importlib.import_module(f"pockit_b200.{scheme or _scheme_of(ref_system)}")
    ref_s = list(ref_system.s)
    S = mod.System([sym.name for sym in ref_s], simplify=bool(getattr(ref_system, "_simplify", False)),
                   fastmath=bool(getattr(ref_system, "_fastmath", False)))
    smap = dict(zip(ref_s, S.s))

    def tr(e):
        return sp.sympify(e).xreplace(smap)

    def bc(v):
        return v if v is None or isinstance(v, (int, float)) else tr(v)

    phases = []
    for p in ref_system.p:
        strip = lambda name: name.rsplit("^{(", 1)[0]  # noqa: E731 -- the reference appends ^{(identifier)}
        q = S.new_phase([strip(sym.name) for sym in p._symbol_state], [strip(sym.name) for sym in p._symbol_control])
        smap.update(zip(p._symbol_state, q.x))
        smap.update(zip(p._symbol_control, q.u))
        smap[p._symbol_time] = q.t
        q.set_dynamics([tr(e) for e in p._expr_dynamics])
        q.set_integral([tr(e) for e in p._expr_integral])
        smap.update(zip(p._symbol_integral, q.I))
        # phase constraints: plain-symbol bounds were split off by the reference; put them back first
        symbols = list(p._symbols)
        n_var = len(p._symbol_state) + len(p._symbol_control)
        cons, lo, hi = [], [], []
        for i, lb, ub in p._variable_bounds_phase:
            cons.append(tr(symbols[i])); lo.append(lb); hi.append(ub)
        for lb, ub in p._time_bounds_phase:
            cons.append(q.t); lo.append(lb); hi.append(ub)
        for i, lb, ub in p._static_parameter_bounds_phase:
            cons.append(tr(symbols[n_var + 1 + i])); lo.append(lb); hi.append(ub)
        for e, lb, ub in zip(p._expr_phase_constraint, p._lower_bound_phase_constraint, p._upper_bound_phase_constraint):
            cons.append(tr(e)); lo.append(float(lb)); hi.append(float(ub))
        q.set_phase_constraint(cons, lo, hi)
        q.set_boundary_condition([bc(v) for v in p._initial_value], [bc(v) for v in p._terminal_value],
                                 bc(p._initial_time), bc(p._terminal_time))
        q.set_discretization(np.asarray(p._mesh, dtype=np.float64), np.asarray(p._num_point, dtype=np.int64))
        phases.append(q)
    S.set_phase(phases)
    S.set_objective(tr(ref_system._expr_objective))
    S.set_system_constraint([tr(c) for c in ref_system._system_constraint_user],
                            list(ref_system._system_constraint_user_lower_bound),
                            list(ref_system._system_constraint_user_upper_bound))
    return S
