"""One phase of an optimal-control problem: modelling front end + lowering of
its discretised callbacks to *segments* the device engine can evaluate.

The user-facing methods mirror ``pockit.base.phasebase.PhaseBase``
(``set_dynamics`` :242, ``set_integral`` :274, ``set_phase_constraint`` :310,
``set_boundary_condition`` :439, ``set_discretization`` :513) with the same
argument meaning and the same ``ValueError`` conditions, so model scripts and
tests written against the reference run unchanged.

Where the reference then builds easyderiv node graphs (:580-626, :661-825) and
NumPy index arrays (:854-995) and, per callback, re-evaluates them with Numba +
fancy indexing (:997-1337), this class emits a flat list of :class:`Segment`
objects -- each one a contiguous run of Jacobian / Hessian slots together with
the product tree (:mod:`pockit_b200.chain`) that produces its values.  The
System turns those into device jobs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Iterable, Optional

import numpy as np
import sympy as sp

from .chain import ONE, GEntry, HEntry, Sym, Term, compose, leaf
from .discretization import Collocation
from .symfunc import SymFunc

__all__ = ["Phase", "BcType", "BcInfo", "Segment"]


class BcType(Enum):
    FREE = 0
    FIXED = 1
    FUNC = 2


@dataclass
class BcInfo:
    t: BcType
    v: None | float | SymFunc


@dataclass
class Segment:
    """A contiguous run of output slots.

    kind      'const'   values are ``data`` (the ±1 of the translation operator)
              'kron'    ``sign * data[a] * [lam[lam_rows[a]]] * term_b`` for a-major (a, b)
              'expand'  ``sign * I.m.data[k] * [lam[lam_off + I.m.row[k]]] * term(node I.m.col[k])``
              'direct'  ``term(node) * [lam[lam_off + node]]``; ``count`` is 1 (front / back) or the
                        number of middle nodes
    nset      which node set evaluates the terms: 'basic' (functions of ``s`` only), 'front', 'mid', 'back'
    """

    kind: str
    count: int
    nset: str = "basic"
    terms: list[Term] = field(default_factory=list)
    data: Optional[np.ndarray] = None
    lam_rows: Optional[np.ndarray] = None  # phase-local constraint rows (kron with multipliers)
    lam_off: int = -1  # phase-local constraint-row offset ('expand' / 'direct' with multipliers)
    sign: float = 1.0
    rows: Optional[np.ndarray] = None  # structure, phase-local (Jacobian: constraint row)
    cols: Optional[np.ndarray] = None
    family: str = ""  # 'dyn' | 'path'


def _affine(base: int, stride: int, count: int) -> np.ndarray:
    if stride == 0:
        return np.full(count, base, dtype=np.int64)
    return base + np.arange(count, dtype=np.int64)


class Phase:
    _scheme: str = ""  # 'lgl' | 'lgr', bound by pockit_b200.lobatto / pockit_b200.radau

    def __init__(
        self,
        identifier: int,
        state: int | list[str],
        control: int | list[str],
        symbol_static_parameter: list[sp.Symbol],
        simplify: bool = False,
        fastmath: bool = False,
    ) -> None:
        def names(spec, stem, what):
            if isinstance(spec, int):
                return [f"{stem}_{i}^{{({identifier})}}" for i in range(spec)]
            if isinstance(spec, list):
                if "t" in spec:
                    raise ValueError(
                        f'Symbol "t" is reserved for time. Use a different name for {what} variables'
                    )
                return [n + f"^{{({identifier})}}" for n in spec]
            raise ValueError(f"{what} must be int or list of str")

        self._identifier = identifier
        self._symbol_state = [sp.Symbol(n) for n in names(state, "x", "state")]
        self._symbol_control = [sp.Symbol(n) for n in names(control, "u", "control")]
        self._symbol_time = sp.Symbol(f"t^{{({identifier})}}")
        self._symbol_static_parameter = symbol_static_parameter
        self._symbols = (
            self._symbol_state + self._symbol_control + [self._symbol_time] + list(symbol_static_parameter)
        )
        self._simplify = simplify
        self._fastmath = fastmath
        self._dynamics_set = self._boundary_condition_set = self._discretization_set = False
        self._func_dynamics: list[SymFunc] = []
        self.col: Optional[Collocation] = None
        self._version = 0  # bumped on every change so Systems know to re-plan
        self.set_integral([])
        self.set_phase_constraint([], [], [])

    # ------------------------------------------------------------------ model API
    def _fn(self, e) -> SymFunc:
        return SymFunc(e, self._symbols, self._simplify)

    def set_dynamics(self, dynamics: list[float | sp.Expr], *, cache: Optional[str] = None):
        if len(dynamics) != self.n_x:
            raise ValueError("the number of dynamics must be equal to the number of state variables")
        self._expr_dynamics = [sp.sympify(d) for d in dynamics]
        self._func_dynamics = [self._fn(d) for d in self._expr_dynamics]
        self._dynamics_set = True
        self._version += 1
        return self

    def set_integral(self, integral: list[float | sp.Expr], *, cache: Optional[str] = None):
        self._expr_integral = [sp.sympify(i) for i in integral]
        self._func_integral = [self._fn(i) for i in self._expr_integral]
        self._symbol_integral = [
            sp.Symbol(f"I_{i}^{{({self._identifier})}}") for i in range(len(self._expr_integral))
        ]
        self._version += 1
        return self

    def set_phase_constraint(
        self,
        phase_constraint: list[sp.Expr],
        lower_bound: list[float],
        upper_bound: list[float],
        bang_bang_control: bool | list[bool] = False,
        *,
        cache: Optional[str] = None,
    ):
        phase_constraint = list(phase_constraint)
        lower_bound = list(lower_bound)
        upper_bound = list(upper_bound)
        if not len(phase_constraint) == len(lower_bound) == len(upper_bound):
            raise ValueError("phase_constraint, lower_bound and upper_bound must have the same length")
        self._variable_bounds_phase = []
        self._static_parameter_bounds_phase = []
        self._time_bounds_phase = []
        self._expr_phase_constraint = []
        lbs, ubs = [], []
        n_var = self.n_x + self.n_u
        for c, lb, ub in zip(phase_constraint, lower_bound, upper_bound):
            c = sp.sympify(c)
            if c.is_symbol:
                # pure symbols become simple bounds (phasebase.py:352-366)
                i = self._symbols.index(c)
                if i < n_var:
                    self._variable_bounds_phase.append((i, lb, ub))
                elif i == n_var:
                    self._time_bounds_phase.append((lb, ub))
                else:
                    self._static_parameter_bounds_phase.append((i - n_var - 1, lb, ub))
            else:
                self._expr_phase_constraint.append(c)
                lbs.append(lb)
                ubs.append(ub)
        self._func_phase_constraint = [self._fn(c) for c in self._expr_phase_constraint]
        self._lower_bound_phase_constraint = np.array(lbs, dtype=np.float64)
        self._upper_bound_phase_constraint = np.array(ubs, dtype=np.float64)
        if isinstance(bang_bang_control, bool):
            flags = [bang_bang_control] * len(phase_constraint)
        else:
            flags = list(bang_bang_control)
        for lb, ub, bb in zip(lower_bound, upper_bound, flags):
            if bb:
                if np.isinf(lb) or np.isinf(ub):
                    raise ValueError(
                        "lower_bound and upper_bound must be finite for bang-bang control constraint"
                    )
                if ub <= lb + 1e-10:
                    raise ValueError(
                        "lower_bound must be strictly less than upper_bound for bang-bang control constraint"
                    )
        self._version += 1
        return self

    def _parse_bc(self, bc) -> BcInfo:
        if bc is None:
            return BcInfo(BcType.FREE, None)
        if isinstance(bc, float):
            return BcInfo(BcType.FIXED, bc)
        if isinstance(bc, sp.Expr):
            return BcInfo(BcType.FUNC, SymFunc(bc, self._symbol_static_parameter, self._simplify))
        raise ValueError("boundary condition must be None, number or sp.Expr")

    def set_boundary_condition(
        self,
        initial_value: list[None | float | sp.Expr],
        terminal_value: list[None | float | sp.Expr],
        initial_time: None | float | sp.Expr,
        terminal_time: None | float | sp.Expr,
        *,
        cache: Optional[str] = None,
    ):
        if not len(initial_value) == len(terminal_value) == self.n_x:
            raise ValueError(
                "initial_value, terminal_value must have the same length as number of state variables"
            )
        as_float = lambda v: float(v) if isinstance(v, (int, np.integer, np.floating)) else v
        initial_value = [as_float(v) for v in initial_value]
        terminal_value = [as_float(v) for v in terminal_value]
        initial_time, terminal_time = as_float(initial_time), as_float(terminal_time)
        self._initial_value, self._terminal_value = initial_value, terminal_value
        self._initial_time, self._terminal_time = initial_time, terminal_time
        self.info_bc_0 = [self._parse_bc(b) for b in initial_value]
        self.info_bc_f = [self._parse_bc(b) for b in terminal_value]
        self.info_t_0 = self._parse_bc(initial_time)
        self.info_t_f = self._parse_bc(terminal_time)
        self._boundary_condition_set = True
        self._version += 1
        return self

    def set_discretization(self, mesh: int | Iterable[float], num_point: int | Iterable[int]):
        if isinstance(mesh, (int, np.integer)):
            if mesh < 1:
                raise ValueError("mesh must contain at least one interval")
            mesh_new = np.linspace(0, 1, int(mesh) + 1, endpoint=True)
        else:
            mesh_new = np.array(list(mesh), dtype=np.float64)
            if mesh_new.ndim != 1 or len(mesh_new) < 2:
                raise ValueError("mesh must contain at least two points")
            if not np.all(np.isfinite(mesh_new)):
                raise ValueError("mesh points must be finite")
            if np.any(np.diff(mesh_new) <= 0):
                raise ValueError("mesh points must be strictly increasing")
            mesh_new = (mesh_new - mesh_new[0]) / (mesh_new[-1] - mesh_new[0])
        n_int = len(mesh_new) - 1
        if isinstance(num_point, (int, np.integer)):
            npt = np.full(n_int, int(num_point), dtype=np.int64)
        else:
            vals = np.array(list(num_point))
            if vals.ndim != 1:
                raise ValueError("num_point must be a one-dimensional iterable")
            if not np.issubdtype(vals.dtype, np.integer):
                raise ValueError("num_point entries must be integers")
            npt = vals.astype(np.int64)
        if len(npt) != n_int:
            raise ValueError("num_point must have the same length as mesh intervals (= len(mesh) - 1)")
        least = 2 if self._scheme == "lgl" else 1
        if np.any(npt < least):
            raise ValueError(f"num_point entries must be at least {least}")
        if np.any(npt > np.iinfo(np.int32).max):
            raise ValueError("num_point entries are too large")
        col = Collocation(self._scheme, mesh_new, npt.astype(np.int32), self.n_x, self.n_u)
        self.col = col
        self._mesh, self._num_interval, self._num_point = mesh_new, n_int, npt.astype(np.int32)
        self._discretization_set = True
        self._version += 1
        return self

    # ------------------------------------------------------------------ properties
    n_x = property(lambda self: len(self._symbol_state))
    n_u = property(lambda self: len(self._symbol_control))
    n = property(lambda self: self.n_x + self.n_u)
    n_s = property(lambda self: len(self._symbol_static_parameter))
    n_I = property(lambda self: len(self._func_integral))
    n_c = property(lambda self: len(self._func_phase_constraint))
    n_d = property(lambda self: self.n_x)
    x = property(lambda self: self._symbol_state)
    u = property(lambda self: self._symbol_control)
    t = property(lambda self: self._symbol_time)
    s = property(lambda self: self._symbol_static_parameter)
    I = property(lambda self: self._symbol_integral)
    F_d = property(lambda self: self._func_dynamics)
    F_I = property(lambda self: self._func_integral)
    F_c = property(lambda self: self._func_phase_constraint)
    c_lb = property(lambda self: self._lower_bound_phase_constraint)
    c_ub = property(lambda self: self._upper_bound_phase_constraint)
    s_b = property(lambda self: self._static_parameter_bounds_phase)
    bc_0 = property(lambda self: self._initial_value)
    bc_f = property(lambda self: self._terminal_value)
    t_0 = property(lambda self: self._initial_time)
    t_f = property(lambda self: self._terminal_time)
    N = property(lambda self: self._num_interval)
    ok = property(
        lambda self: self._dynamics_set and self._boundary_condition_set and self._discretization_set
    )
    l_v = property(lambda self: self.col.l_v)
    r_v = property(lambda self: self.col.r_v)
    l_d = property(lambda self: self.col.l_d)
    r_d = property(lambda self: self.col.r_d)
    l_m = property(lambda self: self.col.l_m)
    r_m = property(lambda self: self.col.r_m)
    t_m = property(lambda self: self.col.t_m)
    w_m = property(lambda self: self.col.w_m)
    L_m = property(lambda self: self.col.L_m)
    L = property(lambda self: self.col.L)
    index_state = property(lambda self: self.col.index_state)
    index_control = property(lambda self: self.col.index_control)
    index_mstage = property(lambda self: self.col.index_mstage)

    @property
    def v_lb(self) -> np.ndarray:
        return self._bounds()[0]

    @property
    def v_ub(self) -> np.ndarray:
        return self._bounds()[1]

    def _bounds(self):
        """Simple bounds on the phase vector (``phasebase.py:632-659``)."""
        lo = np.full(self.L, -np.inf)
        hi = np.full(self.L, np.inf)
        for i, lb, ub in self._variable_bounds_phase:
            sl = slice(self.l_v[i], self.r_v[i])
            lo[sl] = np.maximum(lo[sl], lb)
            hi[sl] = np.minimum(hi[sl], ub)
        for lb, ub in self._time_bounds_phase:
            lo[-2:] = np.maximum(lo[-2:], lb)
            hi[-2:] = np.minimum(hi[-2:], ub)
        return lo, hi

    @staticmethod
    def _value_boundary_condition(info: BcInfo, x, s):
        """Host helper used by pre/post-processing (``phasebase.py:830-837``)."""
        if info.t == BcType.FREE:
            return x
        if info.t == BcType.FIXED:
            return info.v
        f = sp.lambdify(info.v.args, info.v.expr, "math")
        return float(f(*s))

    # ------------------------------------------------------------------ lowering
    def _boundary_sym(self, info: BcInfo, own_index: int, tag: tuple) -> Sym:
        """Lists of a boundary quantity (``_update_node_basic``, phasebase.py:661-724)."""
        if info.t == BcType.FREE:
            return Sym(G=[GEntry(own_index, 0, ONE)])
        if info.t == BcType.FIXED:
            return Sym()
        fn: SymFunc = info.v
        statics = [Sym(G=[GEntry(-self.n_s + k, 0, ONE)]) for k in range(self.n_s)]
        return compose(
            statics,
            fn.G_index,
            [leaf("bG", *tag, jj) for jj in range(fn.n_G)],
            fn.H_index_row,
            fn.H_index_col,
            [leaf("bH", *tag, m) for m in range(fn.n_H)],
        )

    def lower(self) -> "PhaseLowering":
        return PhaseLowering(self)


class PhaseLowering:
    """All symbolic lists and output segments of one phase, for one mesh."""

    def __init__(self, p: Phase):
        self.p = p
        col = p.col
        ms = col.index_mstage
        n_x, n_u, n_s = p.n_x, p.n_u, p.n_s
        self.mid_lo, self.mid_hi = ms.l_m, ms.r_m
        self.n_mid = ms.L_m

        # ---- argument nodes -------------------------------------------------
        statics = [Sym(G=[GEntry(-n_s + k, 0, ONE)]) for k in range(n_s)]
        x_front = [p._boundary_sym(p.info_bc_0[i], int(col.l_v[i]), ("x0", i)) for i in range(n_x)]
        x_back = [p._boundary_sym(p.info_bc_f[i], int(col.r_v[i]) - 1, ("xf", i)) for i in range(n_x)]
        x_mid = [
            Sym(G=[GEntry(int(col.l_v[i]) + col.index_state.l_m, 1, ONE)]) for i in range(n_x)
        ]
        u_front = [
            Sym(G=[GEntry(int(col.l_v[n_x + j]), 0, ONE)]) if ms.f else Sym() for j in range(n_u)
        ]
        u_back = [
            Sym(G=[GEntry(int(col.r_v[n_x + j]) - 1, 0, ONE)]) if ms.b else Sym() for j in range(n_u)
        ]
        u_mid = [
            Sym(G=[GEntry(int(col.l_v[n_x + j]) + col.index_control.l_m, 1, ONE)]) for j in range(n_u)
        ]
        t_front = p._boundary_sym(p.info_t_0, col.L - 2, ("t0",))
        t_back = p._boundary_sym(p.info_t_f, col.L - 1, ("tf",))
        t_mid = compose([t_front, t_back], [0, 1], [leaf("om"), leaf("tm")])
        self.t_delta = compose([t_front, t_back], [0, 1], [Term(None, -1.0), ONE])
        s_mid = [compose([statics[k]], [0], [ONE]) for k in range(n_s)]
        self.x_front, self.x_back = x_front, x_back

        arg_sets = {
            "front": x_front + u_front + [t_front] + statics,
            "mid": x_mid + u_mid + [t_mid] + s_mid,
            "back": x_back + u_back + [t_back] + statics,
        }
        self.sets = ["front"] * ms.f + ["mid"] + ["back"] * ms.b

        def lists_of(fam: str, funcs: list[SymFunc], scaled: bool):
            """Unscaled lists of every function at every node set, and optionally the
            lists of ``dt * f`` (``_update_node_function`` :726-749, ``_update_node_scale`` :751-762)."""
            out = {}
            for nset in self.sets:
                per = []
                for i, fn in enumerate(funcs):
                    un = compose(
                        arg_sets[nset],
                        fn.G_index,
                        [leaf("G", fam, i, jj) for jj in range(fn.n_G)],
                        fn.H_index_row,
                        fn.H_index_col,
                        [leaf("H", fam, i, m) for m in range(fn.n_H)],
                    )
                    if scaled:
                        per.append(
                            compose(
                                [un, self.t_delta], [0, 1], [leaf("dt"), leaf("F", fam, i)], [1], [0], [ONE]
                            )
                        )
                    else:
                        per.append(un)
                out[nset] = per
            return out

        self.dyn = lists_of("d", p.F_d, True)
        self.integ = lists_of("I", p.F_I, True)
        self.path = lists_of("c", p.F_c, False)

    # ------------------------------------------------------------------
    def _count(self, nset: str) -> int:
        return self.n_mid if nset == "mid" else 1

    def jacobian_segments(self) -> list[Segment]:
        """Slot order of ``_update_index_dynamic_constraint`` / ``_update_index_phase_constraint``
        (phasebase.py:854-897, 945-973); values of ``_grad_dynamic_constraint`` /
        ``_grad_phase_constraint`` (:1070-1152)."""
        p, col = self.p, self.p.col
        T, I = col.T, col.I
        ms = col.index_mstage
        segs: list[Segment] = []
        for i in range(p.n_x):
            ld = int(col.l_d[i])
            for part, sym in ((T.f, self.x_front[i]), (None, None), (T.b, self.x_back[i])):
                if part is None:
                    segs.append(
                        Segment(
                            "const",
                            len(T.m),
                            data=T.m.data,
                            rows=ld + T.m.row.astype(np.int64),
                            cols=int(col.l_v[i]) + T.m.col.astype(np.int64),
                            family="dyn",
                        )
                    )
                elif sym.G:
                    nb = len(sym.G)
                    segs.append(
                        Segment(
                            "kron",
                            len(part) * nb,
                            "basic",
                            [g.val for g in sym.G],
                            data=part.data,
                            rows=ld + np.repeat(part.row.astype(np.int64), nb),
                            cols=np.tile(np.array([g.base for g in sym.G], dtype=np.int64), len(part)),
                            family="dyn",
                        )
                    )
        for i in range(p.n_x):
            ld = int(col.l_d[i])
            for nset in self.sets:
                lists = self.dyn[nset][i].G
                if nset == "mid":
                    for g in lists:
                        segs.append(
                            Segment(
                                "expand",
                                len(I.m),
                                "mid",
                                [g.val],
                                sign=-1.0,
                                rows=ld + I.m.row.astype(np.int64),
                                cols=g.base + g.stride * (I.m.col.astype(np.int64) - ms.l_m),
                                family="dyn",
                            )
                        )
                elif lists:
                    part = I.f if nset == "front" else I.b
                    nb = len(lists)
                    segs.append(
                        Segment(
                            "kron",
                            len(part) * nb,
                            nset,
                            [g.val for g in lists],
                            data=part.data,
                            sign=-1.0,
                            rows=ld + np.repeat(part.row.astype(np.int64), nb),
                            cols=np.tile(np.array([g.base for g in lists], dtype=np.int64), len(part)),
                            family="dyn",
                        )
                    )
        r0 = 0
        for q in range(p.n_c):
            for nset in self.sets:
                cnt = self._count(nset)
                first = {"front": 0, "mid": ms.l_m, "back": col.L_m - 1}[nset]
                for g in self.path[nset][q].G:
                    segs.append(
                        Segment(
                            "direct",
                            cnt,
                            nset,
                            [g.val],
                            rows=r0 + first + np.arange(cnt, dtype=np.int64),
                            cols=_affine(g.base, g.stride, cnt),
                            family="path",
                        )
                    )
            r0 += col.L_m
        return segs

    def hessian_segments(self) -> list[Segment]:
        """Slot order of phasebase.py:899-943, 975-995; values of
        ``_hess_dynamic_constraint`` / ``_hess_phase_constraint`` (:1211-1337).
        ``lam_*`` are rows of this phase's constraint block: dynamics rows first,
        then ``n_c * L_m`` path rows."""
        p, col = self.p, self.p.col
        T, I = col.T, col.I
        ms = col.index_mstage
        segs: list[Segment] = []

        def kron_h(part, lists: list[HEntry], ld: int, nset: str, sign: float):
            nb = len(lists)
            return Segment(
                "kron",
                len(part) * nb,
                nset,
                [h.val for h in lists],
                data=part.data,
                lam_rows=ld + part.row.astype(np.int64),
                sign=sign,
                rows=np.tile(np.array([h.row_base for h in lists], dtype=np.int64), len(part)),
                cols=np.tile(np.array([h.col_base for h in lists], dtype=np.int64), len(part)),
                family="dyn",
            )

        for i in range(p.n_x):
            ld = int(col.l_d[i])
            if self.x_front[i].H:
                segs.append(kron_h(T.f, self.x_front[i].H, ld, "basic", 1.0))
            if self.x_back[i].H:
                segs.append(kron_h(T.b, self.x_back[i].H, ld, "basic", 1.0))
        for i in range(p.n_x):
            ld = int(col.l_d[i])
            for nset in self.sets:
                lists = self.dyn[nset][i].H
                if nset == "mid":
                    k = I.m.col.astype(np.int64) - ms.l_m
                    for h in lists:
                        segs.append(
                            Segment(
                                "expand",
                                len(I.m),
                                "mid",
                                [h.val],
                                sign=-1.0,
                                lam_off=ld,
                                rows=h.row_base + h.row_stride * k,
                                cols=h.col_base + h.col_stride * k,
                                family="dyn",
                            )
                        )
                elif lists:
                    segs.append(kron_h(I.f if nset == "front" else I.b, lists, ld, nset, -1.0))
        lam0 = col.n_rows * p.n_x
        for q in range(p.n_c):
            for nset in self.sets:
                cnt = self._count(nset)
                first = {"front": 0, "mid": ms.l_m, "back": col.L_m - 1}[nset]
                for h in self.path[nset][q].H:
                    segs.append(
                        Segment(
                            "direct",
                            cnt,
                            nset,
                            [h.val],
                            lam_off=lam0 + q * col.L_m + first,
                            rows=_affine(h.row_base, h.row_stride, cnt),
                            cols=_affine(h.col_base, h.col_stride, cnt),
                            family="path",
                        )
                    )
        return segs
