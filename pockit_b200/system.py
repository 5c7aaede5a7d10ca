"""Multi-phase NLP assembly behind the Ipopt / SciPy callback contract.

Public surface = ``pockit.base.systembase.SystemBase``: ``new_phase`` (:148),
``set_phase`` (:170), ``set_objective`` (:189), ``set_system_constraint`` (:219),
``update`` (:253) and the callbacks ``objective`` (:602), ``gradient`` (:646),
``constraints`` (:613), ``jacobianstructure`` / ``jacobian`` (:671-693),
``hessianstructure[_o|_c]`` / ``hessian[_o|_c]`` (:726-835), plus the bound and
layout attributes the solver adapters read (``pockit/optimizer/_common.py:9-63``).

Unlike the reference nothing here computes values on the host: the callbacks
hand ``x`` (and multipliers) to the CUDA engine through the C-ABI in
``include/pockit_b200.h`` and raise if that library is missing.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Optional

import numpy as np
import sympy as sp

from .chain import GEntry, HEntry, Term, leaf
from .phase import BcType, Phase, Segment
from .symfunc import SymFunc

__all__ = ["System", "SysSegment", "SystemLowering", "continuous_error_intervals"]


@dataclass
class SysSegment:
    """A run of output slots in a *system-level* array (global indices).

    kind   'phase'   wraps a phase :class:`Segment` (``seg``) of phase ``phase``
           'scaled'  ``(term(node) * w_m[node]) * sys`` -- an integral list seen through a
                     system-level function; ``nset`` / ``phase`` say where ``term`` lives
           'sys'     a single slot holding the system-level leaf ``sys`` itself
           'outer'   ``(A[i] * B[j]) * sys`` for i-major (i, j) over two lists (``pair``)
           'tril'    ``(A[i] * B[j]) * sys`` over the lower triangle i >= j, row-major
    post   None | ('sigma',) | ('lam', i): trailing factor applied last
    """

    kind: str
    count: int
    phase: int = -1
    seg: Optional[Segment] = None
    nset: str = ""
    term: Optional[Term] = None
    sys: Optional[Term] = None
    post: Optional[tuple] = None
    rows: Optional[np.ndarray] = None
    cols: Optional[np.ndarray] = None
    lam_base: int = 0  # global row of the phase's first constraint ('phase' kind)
    pair: Optional[tuple] = None  # ('outer' / 'tril'): the two SysList factors


@dataclass
class SysList:
    """One gradient list of a system-level argument: ``term(node) * w_m[node]`` over
    the nodes of ``nset`` in phase ``phase`` (``term is None``: the constant 1 of a
    static parameter).  ``summed`` marks a broadcast-index list collapsed to the sum
    of its values (easyderiv.py:422-425)."""

    idx: np.ndarray
    count: int
    phase: int = -1
    nset: str = ""
    term: Optional[Term] = None
    summed: bool = False

    def collapsed(self) -> "SysList":
        return SysList(self.idx[:1], 1, self.phase, self.nset, self.term, True)


class System:
    _class_phase: type[Phase] = Phase

    def __init__(self, static_parameter: int | list[str], simplify: bool = False, fastmath: bool = False):
        if isinstance(static_parameter, int):
            names = [f"s_{i}" for i in range(static_parameter)]
        elif isinstance(static_parameter, list):
            names = static_parameter
        else:
            raise ValueError("static_parameter must be int or list of str")
        self._symbol_static_parameter = [sp.Symbol(n) for n in names]
        self._identifier_phase = 0
        self._simplify, self._fastmath = simplify, fastmath
        self._phase: list[Phase] = []
        self._phase_set = self._objective_set = self._system_constraint_set = False
        self._expr_objective = None
        self._system_constraint_user: list = []
        self._system_constraint_user_lower_bound: list = []
        self._system_constraint_user_upper_bound: list = []
        self._lowered: Optional["SystemLowering"] = None
        self._lowered_key = None
        self._engine = None
        # False (default): every callback returns a fresh NumPy array like the reference.
        # True: callbacks return views of engine-owned page-locked buffers (valid until the next
        # call of the same callback) -- what a solver adapter that copies the values anyway wants.
        self.pinned_outputs = False
        self._compact = False
        self.set_phase([])
        self.set_system_constraint([], [], [])

    # ------------------------------------------------------------------ model API
    def new_phase(self, state: int | list[str], control: int | list[str]) -> Phase:
        self._identifier_phase += 1
        return self._class_phase(
            self._identifier_phase - 1,
            state,
            control,
            self._symbol_static_parameter,
            self._simplify,
            self._fastmath,
        )

    def set_phase(self, phase: list[Phase]):
        for i, p in enumerate(phase):
            if not p.ok:
                raise ValueError(
                    f"Dynamics, boundary conditions, "
                    f"or discretization scheme of phase {i} are not fully set"
                )
        self._phase = list(phase)
        self._phase_set = True
        self._invalidate()
        return self

    def set_objective(self, objective: float | sp.Expr, *, cache: Optional[str] = None):
        self._expr_objective = sp.sympify(objective)
        self._objective_set = True
        self._invalidate()
        return self

    def set_system_constraint(
        self,
        system_constraint: list[sp.Expr],
        lower_bound: Iterable[float],
        upper_bound: Iterable[float],
        *,
        cache: Optional[str] = None,
    ):
        lower_bound, upper_bound = list(lower_bound), list(upper_bound)
        if not len(system_constraint) == len(lower_bound) == len(upper_bound):
            raise ValueError("system_constraint, lower_bound and upper_bound must have the same length")
        self._system_constraint_user = list(system_constraint)
        self._system_constraint_user_lower_bound = lower_bound
        self._system_constraint_user_upper_bound = upper_bound
        self._system_constraint_set = True
        self._invalidate()
        return self

    def update(self) -> None:
        """Re-plan after a phase was changed (e.g. re-meshed)."""
        self._invalidate()

    def _invalidate(self):
        self._lowered = None
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    # ------------------------------------------------------------------ layout
    n_s = property(lambda self: len(self._symbol_static_parameter))
    s = property(lambda self: self._symbol_static_parameter)
    n_p = property(lambda self: len(self._phase))
    N = property(lambda self: len(self._phase))
    p = property(lambda self: self._phase)
    ok = property(lambda self: self._phase_set and self._objective_set and self._system_constraint_set)

    @property
    def lowering(self) -> "SystemLowering":
        key = tuple(p._version for p in self._phase)
        if self._lowered is None or key != self._lowered_key:
            self._lowered = SystemLowering(self)
            self._lowered_key = key
        return self._lowered

    l_p = property(lambda self: self.lowering.l_p)
    r_p = property(lambda self: self.lowering.r_p)
    l_i = property(lambda self: self.lowering.l_i)
    r_i = property(lambda self: self.lowering.r_i)
    l_s = property(lambda self: self.lowering.l_s)
    r_s = property(lambda self: self.lowering.r_s)
    L = property(lambda self: self.lowering.r_s)
    n_c = property(lambda self: len(self.lowering.F_c))
    F_o = property(lambda self: self.lowering.F_o)
    F_c = property(lambda self: self.lowering.F_c)
    v_lb = property(lambda self: self.lowering.v_lb)
    v_ub = property(lambda self: self.lowering.v_ub)
    c_lb = property(lambda self: self.lowering.c_lb)
    c_ub = property(lambda self: self.lowering.c_ub)

    # ------------------------------------------------------------------ structures
    @property
    def compact_patterns(self) -> bool:
        """Opt-in, *outside* the reference's pattern contract: ``False`` (default) keeps the
        reference's COO patterns bit for bit (one entry per contributing list element, duplicates
        summed by the consumer, ``optimizer/scipy.py:13-29``).  ``True`` merges duplicate
        ``(row, col)`` pairs: the structures list every pair once (sorted by row, then column), the
        engine sums the duplicates on the device and only the unique values cross PCIe
        (robot_arm LGR 2000x20: Hessian 12.8 M -> 0.56 M values, Jacobian 12.5 M -> 7.9 M).
        ``hessianstructure_o`` / ``_c`` then both return the merged pattern."""
        return self._compact

    @compact_patterns.setter
    def compact_patterns(self, on: bool):
        on = bool(on)
        if on != self._compact:
            self._compact = on
            if self._engine is not None:
                self._apply_compaction(self._engine)

    def _apply_compaction(self, engine):
        from . import plan as P

        for mode, kind in ((P.JAC, "jac"), (P.HESS, "hess")):
            if self._compact:
                c = self.lowering.compaction(kind)
                engine.set_compaction(mode, c["ptr"], c["perm"])
            elif mode in engine.compacted:
                engine.set_compaction(mode, None, None)

    def jacobianstructure(self):
        if self._compact:
            c = self.lowering.compaction("jac")
            return c["row"], c["col"]
        return self.lowering.jac_row, self.lowering.jac_col

    def hessianstructure_o(self):
        if self._compact:
            return self.hessianstructure()
        return self.lowering.hess_o_row, self.lowering.hess_o_col

    def hessianstructure_c(self):
        if self._compact:
            return self.hessianstructure()
        return self.lowering.hess_c_row, self.lowering.hess_c_col

    def hessianstructure(self):
        lo = self.lowering
        if self._compact:
            c = lo.compaction("hess")
            return c["row"], c["col"]
        return (
            np.concatenate([lo.hess_o_row, lo.hess_c_row]),
            np.concatenate([lo.hess_o_col, lo.hess_c_col]),
        )

    # ------------------------------------------------------------------ callbacks (device)
    @property
    def engine(self):
        if self._engine is None or self._engine.lowering is not self.lowering:
            from .engine import Engine  # raises if the CUDA library is unavailable

            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(self.lowering, fastmath=self._fastmath)
            if self._compact:
                self._apply_compaction(self._engine)
        self._engine.reuse_outputs = self.pinned_outputs
        return self._engine

    def evaluate(self, x, fct_c=None, fct_o=1.0):
        """All callbacks at one ``x`` in a single engine call -- what an x-keyed cache in a solver
        adapter asks for (Ipopt evaluates f, grad f, g, J and H at the same ``x``,
        ``optimizer/ipopt.py:41-53``).  ``x`` crosses PCIe once and the copies of the large value
        arrays overlap the remaining compute.  Returns a dict with ``objective``, ``gradient``,
        ``constraints``, ``jacobian`` and, when ``fct_c`` is given, ``hessian``."""
        from . import plan as P

        res = self.engine.evaluate(x, fct_c, fct_o)
        names = {P.OBJ: "objective", P.GRAD: "gradient", P.CONS: "constraints", P.JAC: "jacobian", P.HESS: "hessian"}
        return {names[m]: v for m, v in res.items()}

    # ------------------------------------------------------------------ mesh-refinement data (device)
    def error_estimation_data(self, x):
        """Per phase ``(T_x_aug, I_f_aug)`` at the optimisation vector ``x`` -- what
        ``PhaseBase._error_estimation_data_continuous(x_phase, s)`` (``phasebase.py:1355-1366``)
        returns for every phase: the state interpolant's increments and the integrated dynamics on
        the augmented mesh (one more point per interval)."""
        return self.engine.error_estimation_data(x)

    def check_continuous_intervals(self, x, absolute_tolerance_continuous: float = 1e-8,
                                   relative_tolerance_continuous: float = 1e-8, tolerance_mesh: float = 1e-4):
        """Per phase, which intervals pass the continuous error check at the optimisation vector ``x``."""
        data = self.error_estimation_data(x)
        return [continuous_error_intervals(p, T, I, absolute_tolerance_continuous, relative_tolerance_continuous, tolerance_mesh)
                for p, (T, I) in zip(self._phase, data)]

    def check_continuous(self, value, absolute_tolerance_continuous: float = 1e-8, relative_tolerance_continuous: float = 1e-8,
                         tolerance_mesh: float = 1e-4) -> bool:
        """``True`` if the continuous error of ``value`` is within the tolerance on every interval of
        every phase (``SystemBase.check_continuous``, ``systembase.py:837-900``).  ``value`` is what the
        reference takes -- a ``Variable``, or a list of them followed by the static parameters -- or
        the flat optimisation vector."""
        if isinstance(value, np.ndarray):
            if not self.ok:
                raise ValueError("system is not fully configured")
            x = value
        else:
            from .optimizer._common import pack_guess

            try:
                x, _, _ = pack_guess(self, value, None)
            except ValueError as exc:  # the reference's wording for this entry point
                raise ValueError(str(exc).replace("len(guess)", "len(value)")) from None
        return all(bool(np.all(ok)) for ok in self.check_continuous_intervals(
            x, absolute_tolerance_continuous, relative_tolerance_continuous, tolerance_mesh))

    def check_discontinuous(self, value, *args, **kwargs):
        """Bang-bang (discontinuous) error check: not part of this engine (``phasebase.py:1439-1474``)."""
        raise NotImplementedError("the discontinuous error check is outside the scope of the B200 engine")

    def objective(self, x):
        return self.engine.objective(x)

    def gradient(self, x):
        return self.engine.gradient(x)

    def constraints(self, x):
        return self.engine.constraints(x)

    def jacobian(self, x):
        return self.engine.jacobian(x)

    def hessian_o(self, x):
        return self.engine.hessian_o(x)

    def hessian_c(self, x, fct_c):
        return self.engine.hessian_c(x, fct_c)

    def hessian(self, x, fct_c, fct_o):
        return self.engine.hessian(x, fct_c, fct_o)


def continuous_error_intervals(phase, T_x_aug, I_f_aug, atol: float, rtol: float, mtol: float) -> np.ndarray:
    """Per interval, whether the state increments of the interpolant match the integrated dynamics on
    the augmented mesh (``_error_check_interval_continuous``, ``phasebase.py:1375-1386``).  Intervals
    narrower than ``mtol`` always pass; the column ranges are the reference's ``l_m_aug`` / ``r_m_aug``."""
    from .discretization import AugmentedCollocation

    if getattr(phase, "_aug_cache", (None,))[0] != phase._version:
        phase._aug_cache = (phase._version, AugmentedCollocation(phase.col))
    A = phase._aug_cache[1]
    col = phase.col
    ok = np.ones(len(col.num_point), dtype=bool)
    for k in range(len(ok)):
        if col.mesh[k + 1] - col.mesh[k] < mtol:
            continue
        l, r = int(A.l_m[k]), int(A.r_m[k])
        ok[k] = np.allclose(T_x_aug[:, l:r], I_f_aug[:, l:r], atol=atol, rtol=rtol)
    return ok


def _translate(idx: np.ndarray, l_p: int, r_s: int) -> np.ndarray:
    """Phase-local -> global column (``systembase.py:16-24``): negatives address
    the static parameters at the end of ``x``."""
    idx = np.asarray(idx, dtype=np.int64)
    return np.where(idx >= 0, idx + l_p, idx + r_s)


class SystemLowering:
    """Everything derived from the model at plan time: layout, bounds, the
    COO patterns (``systembase.py:455-551``) and the system-level segments."""

    def __init__(self, system: System):
        self.system = system
        ph = self.phases = list(system._phase)
        n_s = system.n_s
        # --- layout (systembase.py:257-281)
        sizes = np.array([p.L for p in ph], dtype=np.int64)
        self.r_p = np.cumsum(sizes).astype(np.int64)
        self.l_p = self.r_p - sizes
        n_int = np.array([p.n_I for p in ph], dtype=np.int64)
        self.r_i = np.cumsum(n_int).astype(np.int64)
        self.l_i = self.r_i - n_int
        self.l_s = int(self.r_p[-1]) if len(ph) else 0
        self.r_s = self.l_s + n_s
        self.symbols = [sym for p in ph for sym in p.I] + list(system.s)
        self.n_int_total = int(n_int.sum())

        # --- system constraints incl. the ones implied by bounds on FUNC boundaries (:291-364)
        cons = list(system._system_constraint_user)
        lbs = list(system._system_constraint_user_lower_bound)
        ubs = list(system._system_constraint_user_upper_bound)
        for p in ph:
            for i, lb, ub in p._variable_bounds_phase:
                if i < p.n_x and p.info_bc_0[i].t == BcType.FUNC:
                    cons.append(p.bc_0[i]); lbs.append(lb); ubs.append(ub)
                if i < p.n_x and p.info_bc_f[i].t == BcType.FUNC:
                    cons.append(p.bc_f[i]); lbs.append(lb); ubs.append(ub)
            for lb, ub in p._time_bounds_phase:
                if p.info_t_0.t == BcType.FUNC:
                    cons.append(p.t_0); lbs.append(lb); ubs.append(ub)
                if p.info_t_f.t == BcType.FUNC:
                    cons.append(p.t_f); lbs.append(lb); ubs.append(ub)
        static_bounds = []
        exprs, c_lo, c_hi = [], [], []
        for c, lb, ub in zip(cons, lbs, ubs):
            c = sp.sympify(c)
            if c.is_symbol and c in system.s:
                static_bounds.append((system.s.index(c), lb, ub))
            else:
                exprs.append(c); c_lo.append(lb); c_hi.append(ub)
        self.F_c = [SymFunc(c, self.symbols, system._simplify) for c in exprs]
        if system._expr_objective is None:
            raise ValueError("system is not fully configured")
        self.F_o = SymFunc(system._expr_objective, self.symbols, system._simplify)
        n_c = len(self.F_c)

        # which integrals each consumer needs (:413-440)
        def which(free) -> list[np.ndarray]:
            return [np.array([sym in free for sym in p.I], dtype=bool) for p in ph]

        self.which_o = which(self.F_o.free_symbols())
        free_c = set().union(*[f.free_symbols() for f in self.F_c]) if self.F_c else set()
        self.which_c = which(free_c)

        # --- bounds (:553-590)
        s_lo = np.full(n_s, -np.inf)
        s_hi = np.full(n_s, np.inf)
        for p in ph:
            for i, lb, ub in p.s_b:
                s_lo[i] = max(s_lo[i], lb); s_hi[i] = min(s_hi[i], ub)
        for i, lb, ub in static_bounds:
            s_lo[i] = max(s_lo[i], lb); s_hi[i] = min(s_hi[i], ub)
        self.v_lb = np.concatenate([p.v_lb for p in ph] + [s_lo])
        self.v_ub = np.concatenate([p.v_ub for p in ph] + [s_hi])
        lo = [np.array(c_lo, dtype=np.float64)]
        hi = [np.array(c_hi, dtype=np.float64)]
        for p in ph:
            lo += [np.zeros(p.col.n_rows * p.n_x), np.repeat(p.c_lb, p.L_m)]
            hi += [np.zeros(p.col.n_rows * p.n_x), np.repeat(p.c_ub, p.L_m)]
        self.c_lb, self.c_ub = np.concatenate(lo), np.concatenate(hi)
        self.m = len(self.c_lb)

        # --- per-phase lowering and constraint-row offsets
        self.low = [p.lower() for p in ph]
        self.con_base = []  # global row of each phase's first dynamics row
        r = n_c
        for p in ph:
            self.con_base.append(r)
            r += p.col.n_rows * p.n_x + p.n_c * p.L_m

        # --- system-level argument lists: integrals (weights folded in) then statics (:366-411)
        self.integral_lists = []  # per integral: list of (phase, nset, GEntry|HEntry) in reference order
        for pi, (p, lw) in enumerate(zip(ph, self.low)):
            for k in range(p.n_I):
                g = [(pi, ns, e) for ns in lw.sets for e in lw.integ[ns][k].G]
                h = [(pi, ns, e) for ns in lw.sets for e in lw.integ[ns][k].H]
                self.integral_lists.append((g, h))

        self.grad_segments = self._first_order(self.F_o, ("o",))
        self.jac_segments = self._jacobian()
        self.hess_o_segments = self._second_order(self.F_o, ("o",), ("sigma",))
        self.hess_c_segments = self._hessian_c()

        def cat(segs, attr, empty_dtype=np.int32):
            parts = [getattr(s, attr) for s in segs]
            return np.concatenate(parts).astype(np.int64) if parts else np.array([], dtype=empty_dtype)

        self.grad_col = cat(self.grad_segments, "cols")
        self.jac_row, self.jac_col = cat(self.jac_segments, "rows"), cat(self.jac_segments, "cols")
        self.hess_o_row, self.hess_o_col = cat(self.hess_o_segments, "rows"), cat(self.hess_o_segments, "cols")
        self.hess_c_row, self.hess_c_col = cat(self.hess_c_segments, "rows"), cat(self.hess_c_segments, "cols")
        self.nnz_jac = len(self.jac_row)
        self.nnz_hess_o = len(self.hess_o_row)
        self.nnz_hess_c = len(self.hess_c_row)
        self._compaction: dict = {}

    def compaction(self, kind: str) -> dict:
        """De-duplication table of the Jacobian (``'jac'``) or full Hessian (``'hess'``, objective
        part then constraint part) pattern: unique ``(row, col)`` pairs sorted by row then column,
        and for unique entry ``u`` the slots ``perm[ptr[u]:ptr[u+1]]`` (increasing) that carry it."""
        if kind not in self._compaction:
            if kind == "jac":
                row, col = self.jac_row, self.jac_col
            else:
                row = np.concatenate([self.hess_o_row, self.hess_c_row])
                col = np.concatenate([self.hess_o_col, self.hess_c_col])
            key = row.astype(np.int64) * np.int64(max(self.r_s, 1)) + col.astype(np.int64)
            perm = np.argsort(key, kind="stable")
            ks = key[perm]
            start = np.flatnonzero(np.concatenate([[True], ks[1:] != ks[:-1]])) if len(ks) else np.zeros(0, dtype=np.int64)
            ptr = np.concatenate([start, [len(ks)]]).astype(np.int64)
            first = perm[start] if len(ks) else perm
            self._compaction[kind] = dict(row=row[first].astype(np.int64), col=col[first].astype(np.int64), ptr=ptr,
                                          perm=perm.astype(np.int64))
        return self._compaction[kind]

    # ------------------------------------------------------------------
    def _mid_count(self, pi: int, nset: str) -> int:
        return self.low[pi].n_mid if nset == "mid" else 1

    def _gidx(self, pi: int, base: int, stride: int, count: int) -> np.ndarray:
        loc = np.full(count, base, dtype=np.int64) if stride == 0 else base + np.arange(count, dtype=np.int64)
        return _translate(loc, int(self.l_p[pi]), self.r_s)

    def _first_order(self, fn: SymFunc, tag: tuple, row: Optional[int] = None) -> list[SysSegment]:
        """Gradient lists of a system-level function: for every argument with a
        non-zero derivative, the argument's lists times that derivative
        (``forward_gradient_v`` on ``node_objective`` / ``node_system_constraint``,
        systembase.py:646-669; index side :455-498)."""
        out = []
        for jj, j in enumerate(fn.G_index):
            g = leaf("sG", *tag, jj)
            if j < self.n_int_total:
                for pi, nset, e in self.integral_lists[j][0]:
                    cnt = self._mid_count(pi, nset)
                    cols = self._gidx(pi, e.base, e.stride, cnt)
                    out.append(
                        SysSegment(
                            "scaled", cnt, pi, nset=nset, term=e.val, sys=g,
                            rows=None if row is None else np.full(cnt, row, dtype=np.int64), cols=cols,
                        )
                    )
            else:
                k = j - self.n_int_total
                out.append(
                    SysSegment(
                        "sys", 1, sys=g,
                        rows=None if row is None else np.array([row], dtype=np.int64),
                        cols=np.array([self.l_s + k], dtype=np.int64),
                    )
                )
        return out

    def _second_order(self, fn: SymFunc, tag: tuple, post: tuple) -> list[SysSegment]:
        """Hessian lists of a system-level function (``forward_hessian_system_v``,
        easyderiv.py:433-459; index side :358-390)."""
        out = []
        for jj, j in enumerate(fn.G_index):
            if j >= self.n_int_total:
                continue  # static parameters have no second-order lists of their own
            g = leaf("sG", *tag, jj)
            for pi, nset, e in self.integral_lists[j][1]:
                cnt = self._mid_count(pi, nset)
                out.append(
                    SysSegment(
                        "scaled", cnt, pi, nset=nset, term=e.val, sys=g, post=post,
                        rows=self._gidx(pi, e.row_base, e.row_stride, cnt),
                        cols=self._gidx(pi, e.col_base, e.col_stride, cnt),
                    )
                )
        for m, (r, c) in enumerate(zip(fn.H_index_row, fn.H_index_col)):
            h = leaf("sH", *tag, m)
            diag = r == c
            for a in self._arg_lists(r):
                for b in self._arg_lists(c):
                    # easyderiv.py:408-429 (values) / :330-354 (indices)
                    if a.idx[0] < b.idx[0]:
                        if diag:
                            continue
                        row, col = b, a
                    else:
                        row, col = a, b
                    if row.idx[0] > col.idx[0]:
                        out.append(
                            SysSegment(
                                "outer", row.count * col.count, sys=h, post=post, pair=(row, col),
                                rows=np.repeat(row.idx, col.count), cols=np.tile(col.idx, row.count),
                            )
                        )
                        continue
                    if row.count > 1 and row.idx[0] == row.idx[-1]:
                        row = row.collapsed()
                    if col.count > 1 and col.idx[0] == col.idx[-1]:
                        col = col.collapsed()
                    tr, tc = np.tril_indices(row.count)
                    for first, second in ((row, col),) if diag else ((row, col), (col, row)):
                        out.append(
                            SysSegment(
                                "tril", len(tr), sys=h, post=post, pair=(first, second),
                                rows=row.idx[tr], cols=row.idx[tc],
                            )
                        )
        return out

    def _arg_lists(self, j: int) -> list["SysList"]:
        """Global gradient lists of system-level argument ``j`` (an integral, weights
        folded in -- ``_translate_value`` systembase.py:27-47 -- or a static parameter)."""
        if j >= self.n_int_total:
            k = j - self.n_int_total
            return [SysList(np.array([self.l_s + k], dtype=np.int64), 1)]
        out = []
        for pi, nset, e in self.integral_lists[j][0]:
            cnt = self._mid_count(pi, nset)
            if cnt == 0:
                # a phase without middle nodes (one interval of minimal order): the list is empty and
                # contributes no slots whatever it is paired with (the reference reads a_i[0] of the
                # empty array here, easyderiv.py:333, and then emits zero-length index arrays)
                continue
            out.append(SysList(self._gidx(pi, e.base, e.stride, cnt), cnt, pi, nset, e.val))
        return out

    def _jacobian(self) -> list[SysSegment]:
        segs = []
        for i, fn in enumerate(self.F_c):
            segs += self._first_order(fn, ("c", i), row=i)
        for pi, (p, lw) in enumerate(zip(self.phases, self.low)):
            base = self.con_base[pi]
            path_base = base + p.col.n_rows * p.n_x
            for sg in lw.jacobian_segments():
                off = base if sg.family == "dyn" else path_base
                segs.append(
                    SysSegment(
                        "phase", sg.count, pi, seg=sg, lam_base=base,
                        rows=sg.rows + off, cols=_translate(sg.cols, int(self.l_p[pi]), self.r_s),
                    )
                )
        return segs

    def _hessian_c(self) -> list[SysSegment]:
        segs = []
        for i, fn in enumerate(self.F_c):
            segs += self._second_order(fn, ("c", i), ("lam", i))
        for pi, (p, lw) in enumerate(zip(self.phases, self.low)):
            base = self.con_base[pi]
            for sg in lw.hessian_segments():
                segs.append(
                    SysSegment(
                        "phase", sg.count, pi, seg=sg, lam_base=base,
                        rows=_translate(sg.rows, int(self.l_p[pi]), self.r_s),
                        cols=_translate(sg.cols, int(self.l_p[pi]), self.r_s),
                    )
                )
        return segs
