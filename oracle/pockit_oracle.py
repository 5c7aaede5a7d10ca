"""CPU oracle for the NLP-callback hot path -- TEST INFRASTRUCTURE ONLY.

A plain NumPy / SciPy restatement of the reference's algorithm
(``pockit/base/phasebase.py:839-1337``, ``pockit/base/systembase.py:592-835``,
``pockit/base/easyderiv.py:97-459``, ``pockit/base/fastfunc.py:237-296``): every
derivative list is a NumPy array, lists are combined by explicit chain-rule
passes, Jacobian / Hessian values are expanded through the integration
operator with fancy indexing, and everything is concatenated on the host --
deliberately the *array-at-a-time* formulation of the reference, not the
product's fused formulation, so that the two can check each other.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product package never
does.  It consumes a ``pockit_b200`` System purely as a *model description*
(SymPy expressions, boundary-condition kinds, mesh tables).

Parity pin: ``tests/golden/*.npz`` hold outputs of the real reference
(``/root/reference`` imported in the build container by
``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks this
module against them -- structures bit-exact, values to 1e-12 rel / 1e-14 abs.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse
import sympy as sp
from sympy.codegen.rewriting import create_expand_pow_optimization

_expand_pow = create_expand_pow_optimization(3)


# ---------------------------------------------------------------------------
# vectorised function + sparse symbolic derivatives   (fastfunc.py:237-296)
# ---------------------------------------------------------------------------
class OFunc:
    def __init__(self, symfunc):
        self.G_index = np.asarray(symfunc.G_index, dtype=np.int32)
        self.H_index_row = np.asarray(symfunc.H_index_row, dtype=np.int32)
        self.H_index_col = np.asarray(symfunc.H_index_col, dtype=np.int32)
        args = symfunc.args
        self.n_args = len(args)

        def build(exprs):
            if not exprs:
                return None
            # one CSE per output group with the 'basic' pre/post passes, x**2 / x**3
            # expanded to products -- the evaluation order of fastfunc.py:271-296
            cse = lambda e: sp.cse(e, optimizations="basic")
            return sp.lambdify(args, [_expand_pow(e) for e in exprs], modules="numpy", cse=cse)

        self._F = build([symfunc.expr])
        self._G = build(symfunc.G_expr)
        self._H = build(symfunc.H_expr)

    def _call(self, fn, count, x, n):
        out = np.empty((count, n), dtype=np.float64)
        if fn is None:
            return out
        cols = [x[i * n : (i + 1) * n] for i in range(self.n_args)]
        with np.errstate(all="ignore"):
            vals = fn(*cols)
        for r, v in enumerate(vals):
            out[r] = v
        return out

    def F(self, x, n):
        return self._call(self._F, 1, x, n)[0]

    def G(self, x, n):
        return self._call(self._G, len(self.G_index), x, n)

    def H(self, x, n):
        return self._call(self._H, len(self.H_index_row), x, n)


# ---------------------------------------------------------------------------
# chain-rule graph   (easyderiv.py)
# ---------------------------------------------------------------------------
def _less(a, b):  # easyderiv.py:8-19
    if a < 0:
        return b < 0 and a < b
    return b < 0 or a < b


class ONode:
    def __init__(self, l=1, args=None):
        self.l = l
        self.args = args or []
        self.g = np.empty((0, 1))
        self.g_i = np.empty(0, dtype=np.int32)
        self.h = np.empty((0, 1))
        self.h_i_row = np.empty(0, dtype=np.int32)
        self.h_i_col = np.empty(0, dtype=np.int32)
        self.G, self.G_i = [], []
        self.H, self.H_i_row, self.H_i_col = [], [], []

    def leaf(self, index, value=None):
        index = np.atleast_1d(np.asarray(index, dtype=np.int64))
        self.G_i = [index]
        self.G = [np.full(len(index), 1.0) if value is None else value]
        return self

    def bind(self, fn: OFunc):
        self.g_i, self.h_i_row, self.h_i_col = fn.G_index, fn.H_index_row, fn.H_index_col
        return self


def _wide(a, l):
    return np.full(l, a[0], dtype=a.dtype) if len(a) == 1 and l > 1 else a


def gradient_indices(nodes):  # easyderiv.py:97-117
    for nd in nodes:
        if len(nd.g_i):
            nd.G_i = [_wide(a, nd.l) for j in nd.g_i for a in nd.args[j].G_i]


def gradient_values(nodes):  # easyderiv.py:120-140
    for nd in nodes:
        if len(nd.g_i):
            # zip() in the reference stops at the shorter of (args, local values): a node whose
            # local values were not refreshed (integral not needed by the caller) yields no lists
            n = min(len(nd.g_i), len(nd.g))
            nd.G = [_wide(a, nd.l) * nd.g[jj] for jj, j in enumerate(nd.g_i[:n]) for a in nd.args[j].G]


def hessian_indices_phase(nodes):  # easyderiv.py:143-228
    for nd in nodes:
        if not nd.args:
            continue
        rows, cols = [], []
        for j in nd.g_i:
            for r_, c_ in zip(nd.args[j].H_i_row, nd.args[j].H_i_col):
                rows.append(_wide(r_, nd.l))
                cols.append(_wide(c_, nd.l))
        for hr, hc in zip(nd.h_i_row, nd.h_i_col):
            for a in nd.args[hr].G_i:
                for b in nd.args[hc].G_i:
                    a2, b2 = _wide(a, nd.l), _wide(b, nd.l)
                    if len(a) == 1 and len(b) > 1:
                        a2 = np.full(len(b), a[0], dtype=a.dtype)
                    if len(b) == 1 and len(a) > 1:
                        b2 = np.full(len(a), b[0], dtype=b.dtype)
                    if hr == hc:
                        if not _less(a[0], b[0]):
                            rows.append(a2)
                            cols.append(b2)
                    elif _less(a[0], b[0]):
                        rows.append(b2)
                        cols.append(a2)
                    else:
                        rows.append(a2)
                        cols.append(b2)
        nd.H_i_row, nd.H_i_col = rows, cols


def hessian_values_phase(nodes):  # easyderiv.py:231-304
    for nd in nodes:
        if not nd.args:
            continue
        n = min(len(nd.g_i), len(nd.g))
        out = [_wide(a, nd.l) * nd.g[jj] for jj, j in enumerate(nd.g_i[:n]) for a in nd.args[j].H]
        for m, (hr, hc) in enumerate(zip(nd.h_i_row[: len(nd.h)], nd.h_i_col)):
            for ai, av in zip(nd.args[hr].G_i, nd.args[hr].G):
                for bi, bv in zip(nd.args[hc].G_i, nd.args[hc].G):
                    av2, bv2 = _wide(av, nd.l), _wide(bv, nd.l)
                    if hr == hc:
                        if not _less(ai[0], bi[0]):
                            out.append(av2 * bv2 * nd.h[m])
                    elif ai[0] == bi[0]:
                        out.append(av2 * bv2 * nd.h[m] * 2)
                    else:
                        out.append(av2 * bv2 * nd.h[m])
        nd.H = out


def _system_pairs(nd):
    """Shared traversal of easyderiv.py:323-355 (indices) and :393-430 (values):
    yields ``(row_i, row_v, col_i, col_v, h, diag)`` after the lower-triangle swap."""
    for m, (hr, hc) in enumerate(zip(nd.h_i_row, nd.h_i_col)):
        diag = hr == hc
        A, B = nd.args[hr], nd.args[hc]
        for ai, av in zip(A.G_i, A.G if A.G else [None] * len(A.G_i)):
            for bi, bv in zip(B.G_i, B.G if B.G else [None] * len(B.G_i)):
                if ai[0] < bi[0]:
                    if diag:
                        continue
                    yield bi, bv, ai, av, m, diag
                else:
                    yield ai, av, bi, bv, m, diag


def hessian_indices_system(nodes):  # easyderiv.py:323-390
    for nd in nodes:
        if not nd.args:
            continue
        rows, cols = [], []
        for j in nd.g_i:
            for r_, c_ in zip(nd.args[j].H_i_row, nd.args[j].H_i_col):
                rows.append(_wide(r_, nd.l))
                cols.append(_wide(c_, nd.l))
        for ri, _, ci, _, m, diag in _system_pairs(nd):
            if ri[0] > ci[0]:
                rows.append(np.repeat(ri, len(ci)))
                cols.append(np.tile(ci, len(ri)))
            else:
                if len(ri) > 1 and ri[0] == ri[-1]:
                    ri = ri[:1]
                tr, tc = np.tril_indices(len(ri))
                for _ in range(1 if diag else 2):
                    rows.append(ri[tr])
                    cols.append(ri[tc])
        nd.H_i_row, nd.H_i_col = rows, cols


def hessian_values_system(nodes):  # easyderiv.py:393-459
    for nd in nodes:
        if not nd.args:
            continue
        out = [_wide(a, nd.l) * nd.g[jj] for jj, j in enumerate(nd.g_i) for a in nd.args[j].H]
        for ri, rv, ci, cv, m, diag in _system_pairs(nd):
            h = nd.h[m]
            if ri[0] > ci[0]:
                out.append(np.kron(rv, cv) * h)
            else:
                if len(ri) > 1 and ri[0] == ri[-1]:
                    rv = np.array([np.sum(rv)])
                if len(ci) > 1 and ci[0] == ci[-1]:
                    cv = np.array([np.sum(cv)])
                tr, tc = np.tril_indices(len(rv))
                out.append(rv[tr] * cv[tc] * h)
                if not diag:
                    out.append(cv[tr] * rv[tc] * h)
        nd.H = out


def _cat(parts, dtype=np.float64):
    parts = list(parts)
    return np.concatenate(parts).astype(dtype) if parts else np.array([], dtype=dtype)


# ---------------------------------------------------------------------------
# one phase   (phasebase.py)
# ---------------------------------------------------------------------------
FREE, FIXED, FUNC = 0, 1, 2


class OPhase:
    def __init__(self, p):
        self.col = col = p.col
        self.n_x, self.n_u, self.n_s = p.n_x, p.n_u, p.n_s
        self.L, self.L_m = col.L, col.L_m
        self.ms = col.index_mstage
        self.F_d = [OFunc(f) for f in p.F_d]
        self.F_I = [OFunc(f) for f in p.F_I]
        self.F_c = [OFunc(f) for f in p.F_c]
        self.n_I, self.n_c = len(self.F_I), len(self.F_c)

        def bc(info):
            kind = info.t.value
            return (kind, OFunc(info.v) if kind == FUNC else info.v)

        self.bc_0 = [bc(i) for i in p.info_bc_0]
        self.bc_f = [bc(i) for i in p.info_bc_f]
        self.bc_t0, self.bc_tf = bc(p.info_t_0), bc(p.info_t_f)

        # sparse operators, same storage the reference multiplies with (SciPy CSR)
        def csr(op):
            r = np.concatenate([op.f.row, op.m.row, op.b.row])
            c = np.concatenate([op.f.col, op.m.col, op.b.col])
            d = np.concatenate([op.f.data, op.m.data, op.b.data])
            return scipy.sparse.coo_array((d, (r, c)), shape=op.shape).tocsr()

        self.T_csr, self.I_csr = csr(col.T), csr(col.I)
        self._build_nodes()
        self._build_indices()

    # -- graph construction (phasebase.py:125-194, 580-626, 661-825)
    def _build_nodes(self):
        col, ms, n_x, n_u, n_s = self.col, self.ms, self.n_x, self.n_u, self.n_s
        Lmid = ms.L_m
        self.static = [ONode().leaf(-n_s + k) for k in range(n_s)]

        def boundary(kind_v, own):
            kind, v = kind_v
            nd = ONode()
            if kind == FREE:
                nd.leaf(own)
            elif kind == FUNC:
                nd.args = self.static
                nd.bind(v)
            return nd

        self.x_front = [boundary(self.bc_0[i], col.l_v[i]) for i in range(n_x)]
        self.x_back = [boundary(self.bc_f[i], col.r_v[i] - 1) for i in range(n_x)]
        self.x_mid = [
            ONode(Lmid).leaf(np.arange(col.l_v[i] + col.index_state.l_m, col.l_v[i] + col.index_state.r_m))
            for i in range(n_x)
        ]
        self.u_front = [ONode().leaf(col.l_v[n_x + j]) if ms.f else ONode() for j in range(n_u)]
        self.u_back = [ONode().leaf(col.r_v[n_x + j] - 1) if ms.b else ONode() for j in range(n_u)]
        self.u_mid = [
            ONode(Lmid).leaf(
                np.arange(col.l_v[n_x + j] + col.index_control.l_m, col.l_v[n_x + j] + col.index_control.r_m)
            )
            for j in range(n_u)
        ]
        self.t_front = boundary(self.bc_t0, self.L - 2)
        self.t_back = boundary(self.bc_tf, self.L - 1)
        self.t_mid = ONode(Lmid, [self.t_front, self.t_back])
        self.t_mid.g_i = np.array([0, 1])
        self.t_mid.g = np.array([1.0 - col.t_m[ms.m], col.t_m[ms.m]])
        self.t_delta = ONode(1, [self.t_front, self.t_back])
        self.t_delta.g_i = np.array([0, 1])
        self.t_delta.g = np.array([[-1.0], [1.0]])
        self.s_mid = []
        for k in range(n_s):
            nd = ONode(Lmid, [self.static[k]])
            nd.g_i = np.array([0])
            nd.g = np.array([np.full(Lmid, 1.0)])
            self.s_mid.append(nd)
        self.basic = (
            self.x_front + self.x_back + [self.t_front, self.t_back, self.t_mid, self.t_delta] + self.s_mid
        )
        gradient_indices(self.basic)
        hessian_indices_phase(self.basic)
        self.arg_front = self.x_front + self.u_front + [self.t_front] + self.static
        self.arg_mid = self.x_mid + self.u_mid + [self.t_mid] + self.s_mid
        self.arg_back = self.x_back + self.u_back + [self.t_back] + self.static

        def family(funcs, scaled):
            un = {
                "f": [ONode(1, self.arg_front).bind(f) for f in funcs],
                "m": [ONode(Lmid, self.arg_mid).bind(f) for f in funcs],
                "b": [ONode(1, self.arg_back).bind(f) for f in funcs],
            }
            order = un["f"] + un["m"] + un["b"]
            sc = None
            if scaled:
                sc = {}
                for key, nodes in un.items():
                    sc[key] = []
                    for nd in nodes:
                        s_ = ONode(nd.l, [nd, self.t_delta])
                        s_.g_i = np.array([0, 1])
                        s_.h_i_row, s_.h_i_col = np.array([1]), np.array([0])
                        s_.h = np.array([[1.0]])
                        sc[key].append(s_)
                order = order + sc["f"] + sc["m"] + sc["b"]
            gradient_indices(order)
            hessian_indices_phase(order)
            return un, sc, order

        self.dyn_u, self.dyn, self.dyn_all = family(self.F_d, True)
        self.int_u, self.int, self.int_all = family(self.F_I, True)
        self.pc, _, self.pc_all = family(self.F_c, False)

    # -- index arrays (phasebase.py:854-995)
    def _build_indices(self):
        col, ms = self.col, self.ms
        T, I = col.T, col.I
        jr, jc = [], []
        for i in range(self.n_x):
            for part, nd in ((T.f, self.x_front[i]), (T.m, None), (T.b, self.x_back[i])):
                if nd is None:
                    jr.append(col.l_d[i] + part.row)
                    jc.append(col.l_v[i] + part.col)
                elif nd.G_i:
                    gi = np.concatenate(nd.G_i)
                    jr.append(col.l_d[i] + np.repeat(part.row, len(gi)))
                    jc.append(np.tile(gi, len(part)))
        for i in range(self.n_x):
            if ms.f and self.dyn["f"][i].G_i:
                gi = np.concatenate(self.dyn["f"][i].G_i)
                jr.append(col.l_d[i] + np.repeat(I.f.row, len(gi)))
                jc.append(np.tile(gi, len(I.f)))
            for gi in self.dyn["m"][i].G_i:
                jr.append(col.l_d[i] + I.m.row)
                jc.append(gi[I.m.col - ms.l_m])
            if ms.b and self.dyn["b"][i].G_i:
                gi = np.concatenate(self.dyn["b"][i].G_i)
                jr.append(col.l_d[i] + np.repeat(I.b.row, len(gi)))
                jc.append(np.tile(gi, len(I.b)))
        self.jac_dyn_row, self.jac_dyn_col = _cat(jr, np.int64), _cat(jc, np.int64)

        hr, hc = [], []
        for i in range(self.n_x):
            for part, nd in ((T.f, self.x_front[i]), (T.b, self.x_back[i])):
                if nd.H_i_row:
                    hr.append(np.tile(np.concatenate(nd.H_i_row), len(part)))
                    hc.append(np.tile(np.concatenate(nd.H_i_col), len(part)))
        for i in range(self.n_x):
            if ms.f and self.dyn["f"][i].H_i_row:
                hr.append(np.tile(np.concatenate(self.dyn["f"][i].H_i_row), len(I.f)))
                hc.append(np.tile(np.concatenate(self.dyn["f"][i].H_i_col), len(I.f)))
            for r_, c_ in zip(self.dyn["m"][i].H_i_row, self.dyn["m"][i].H_i_col):
                hr.append(r_[I.m.col - ms.l_m])
                hc.append(c_[I.m.col - ms.l_m])
            if ms.b and self.dyn["b"][i].H_i_row:
                hr.append(np.tile(np.concatenate(self.dyn["b"][i].H_i_row), len(I.b)))
                hc.append(np.tile(np.concatenate(self.dyn["b"][i].H_i_col), len(I.b)))
        self.hess_dyn_row, self.hess_dyn_col = _cat(hr, np.int64), _cat(hc, np.int64)

        jr, jc, hr, hc = [], [], [], []
        r_ = 0
        for q in range(self.n_c):
            if ms.f:
                for gi in self.pc["f"][q].G_i:
                    jr.append(np.array([r_]))
                    jc.append(gi)
            for gi in self.pc["m"][q].G_i:
                jr.append(np.arange(r_ + ms.l_m, r_ + ms.r_m))
                jc.append(gi)
            if ms.b:
                for gi in self.pc["b"][q].G_i:
                    jr.append(np.array([r_ + self.L_m - 1]))
                    jc.append(gi)
            r_ += self.L_m
            for key, on in (("f", ms.f), ("m", True), ("b", ms.b)):
                if on:
                    hr += self.pc[key][q].H_i_row
                    hc += self.pc[key][q].H_i_col
        self.jac_pc_row, self.jac_pc_col = _cat(jr, np.int64), _cat(jc, np.int64)
        self.hess_pc_row, self.hess_pc_col = _cat(hr, np.int64), _cat(hc, np.int64)

    # -- values
    @staticmethod
    def _bc_value(kind_v, x, s):  # phasebase.py:830-837
        kind, v = kind_v
        if kind == FREE:
            return x
        if kind == FIXED:
            return v
        return v.F(s, 1)[0]

    def value_basic(self, x, s):  # phasebase.py:839-852 (works on a copy, never the caller's x)
        col = self.col
        x = x.copy()
        for i in range(self.n_x):
            x[col.l_v[i]] = self._bc_value(self.bc_0[i], x[col.l_v[i]], s)
            x[col.r_v[i] - 1] = self._bc_value(self.bc_f[i], x[col.r_v[i] - 1], s)
        x[-2] = self._bc_value(self.bc_t0, x[-2], s)
        x[-1] = self._bc_value(self.bc_tf, x[-1], s)
        mt = (x[-1] + x[-2]) / 2
        dt = x[-1] - x[-2]
        t_ = (col.t_m - 0.5) * dt + mt
        if col.scheme == "lgl":
            mid = x[:-2]
        else:  # drop every state's terminal node (radau/discretization.py:143-166)
            mid = np.concatenate(
                [x[col.l_v[i] : col.r_v[i] - 1] for i in range(self.n_x)]
                + [x[col.l_v[self.n_x + j] : col.r_v[self.n_x + j]] for j in range(self.n_u)]
            )
        return np.concatenate([mid, t_, np.repeat(s, self.L_m)]), dt, x

    def value_integral(self, which, x, s):  # :997-1006
        vb, dt, _ = self.value_basic(x, s)
        return np.array(
            [self.F_I[k].F(vb, self.L_m).dot(self.col.w_m) * dt if flag else 0.0 for k, flag in enumerate(which)]
        )

    def value_dynamic(self, x, s):  # :1008-1012
        vb, dt, xs = self.value_basic(x, s)
        col = self.col
        out = []
        for i in range(self.n_x):
            tx = self.T_csr.dot(xs[col.l_v[i] : col.r_v[i]])
            out.append(tx - self.I_csr.dot(self.F_d[i].F(vb, self.L_m)) * dt)
        return _cat(out)

    def error_estimation_data_continuous(self, x, s):  # :1339-1366
        """``(T_x_aug . x, dt * I_m_aug . f(V_xu_aug . x))`` on the augmented mesh, each ``[n_x][rows]``."""
        from pockit_b200.discretization import AugmentedCollocation

        if not hasattr(self, "_aug"):
            self._aug = AugmentedCollocation(self.col)
        A, col = self._aug, self.col
        _, dt, xs = self.value_basic(x, s)  # boundary values substituted first (:1340-1347)
        mt = (xs[-1] + xs[-2]) / 2
        t_aug = (A.t_m - 0.5) * dt + mt
        xu_aug = A.V.dot(xs[: col.L_xu])
        vb_aug = np.concatenate([xu_aug, t_aug, np.repeat(s, A.L_m)])
        L_x_all = col.r_v[self.n_x - 1] if self.n_x else 0
        T_x = A.T.dot(xs[:L_x_all]).reshape(self.n_x, -1)
        I_f = np.array([A.I.dot(f.F(vb_aug, A.L_m)) for f in self.F_d]).reshape(self.n_x, -1) * dt
        return T_x, I_f

    def value_path(self, x, s):  # :1014-1021
        vb, _, _ = self.value_basic(x, s)
        return _cat([f.F(vb, self.L_m) for f in self.F_c])

    def _refresh_basic(self, s, second):  # :1023-1034 / :1154-1170
        for info, nd in list(zip(self.bc_0, self.x_front)) + list(zip(self.bc_f, self.x_back)) + [
            (self.bc_t0, self.t_front), (self.bc_tf, self.t_back)
        ]:
            if info[0] == FUNC:
                nd.g = info[1].G(s, 1)
                if second:
                    nd.h = info[1].H(s, 1)
        gradient_values(self.basic)
        if second:
            hessian_values_phase(self.basic)

    def _load(self, funcs, un, sc, vb, dt, second, which=None):
        """Local derivative values of every function at front / middle / back nodes
        (the repeated blocks of phasebase.py:1036-1068, 1083-1113, 1130-1139, 1172-1267)."""
        ms = self.ms
        for i, fn in enumerate(funcs):
            if which is not None and not which[i]:
                continue
            f = fn.F(vb, self.L_m)
            g = fn.G(vb, self.L_m)
            h = fn.H(vb, self.L_m) if second else None
            for key, sl, on in (("f", slice(0, 1), ms.f), ("m", ms.m, True), ("b", slice(-1, None), ms.b)):
                if not on:
                    continue
                un[key][i].g = g[:, sl]
                if second:
                    un[key][i].h = h[:, sl]
                if sc is not None:
                    fv = f[sl]
                    sc[key][i].g = np.array([np.full_like(fv, dt), fv])

    def grad_dynamic(self, x, s):  # :1070-1128
        col, ms = self.col, self.ms
        T, I = col.T, col.I
        out = []
        for i in range(self.n_x):
            if self.x_front[i].G:
                out.append(np.kron(T.f.data, np.concatenate(self.x_front[i].G)))
            out.append(T.m.data)
            if self.x_back[i].G:
                out.append(np.kron(T.b.data, np.concatenate(self.x_back[i].G)))
        vb, dt, _ = self.value_basic(x, s)
        self._load(self.F_d, self.dyn_u, self.dyn, vb, dt, False)
        gradient_values(self.dyn_all)
        for i in range(self.n_x):
            if ms.f and self.dyn["f"][i].G:
                out.append(-np.kron(I.f.data, np.concatenate(self.dyn["f"][i].G)))
            for G_ in self.dyn["m"][i].G:
                out.append(-I.m.data * G_[I.m.col - ms.l_m])
            if ms.b and self.dyn["b"][i].G:
                out.append(-np.kron(I.b.data, np.concatenate(self.dyn["b"][i].G)))
        return _cat(out)

    def grad_path(self, x, s):  # :1130-1152
        vb, dt, _ = self.value_basic(x, s)
        self._load(self.F_c, self.pc, None, vb, dt, False)
        gradient_values(self.pc_all)
        ms = self.ms
        out = []
        for q in range(self.n_c):
            for key, on in (("f", ms.f), ("m", True), ("b", ms.b)):
                if on:
                    out += self.pc[key][q].G
        return _cat(out)

    def hess_dynamic(self, x, s, lam):  # :1211-1301
        col, ms = self.col, self.ms
        T, I = col.T, col.I
        out = []
        for i in range(self.n_x):
            for part, nd in ((T.f, self.x_front[i]), (T.b, self.x_back[i])):
                if nd.H:
                    out.append(np.kron(part.data * lam[col.l_d[i] + part.row], np.concatenate(nd.H)))
        vb, dt, _ = self.value_basic(x, s)
        self._load(self.F_d, self.dyn_u, self.dyn, vb, dt, True)
        gradient_values(self.dyn_all)
        hessian_values_phase(self.dyn_all)
        for i in range(self.n_x):
            if ms.f and self.dyn["f"][i].H:
                out.append(-np.kron(I.f.data * lam[col.l_d[i] + I.f.row], np.concatenate(self.dyn["f"][i].H)))
            for H_ in self.dyn["m"][i].H:
                out.append(-I.m.data * lam[col.l_d[i] + I.m.row] * H_[I.m.col - ms.l_m])
            if ms.b and self.dyn["b"][i].H:
                out.append(-np.kron(I.b.data * lam[col.l_d[i] + I.b.row], np.concatenate(self.dyn["b"][i].H)))
        return _cat(out)

    def hess_path(self, x, s, lam):  # :1303-1337
        vb, dt, _ = self.value_basic(x, s)
        self._load(self.F_c, self.pc, None, vb, dt, True)
        gradient_values(self.pc_all)
        hessian_values_phase(self.pc_all)
        ms = self.ms
        out = []
        f_ = 0
        for q in range(self.n_c):
            if ms.f:
                out += [H_ * lam[f_] for H_ in self.pc["f"][q].H]
            out += [H_ * lam[f_ + ms.l_m : f_ + ms.r_m] for H_ in self.pc["m"][q].H]
            if ms.b:
                out += [H_ * lam[f_ + self.L_m - 1] for H_ in self.pc["b"][q].H]
            f_ += self.L_m
        return _cat(out)

    def integral_lists(self, which, x, s, second):  # :1036-1068 / :1172-1209
        self._refresh_basic(s, second)
        vb, dt, _ = self.value_basic(x, s)
        self._load(self.F_I, self.int_u, self.int, vb, dt, second, which)
        gradient_values(self.int_all)
        if second:
            hessian_values_phase(self.int_all)


# ---------------------------------------------------------------------------
# the system   (systembase.py)
# ---------------------------------------------------------------------------
def _translate(idx, l_p, r_s):  # systembase.py:16-24
    idx = np.asarray(idx, dtype=np.int64)
    return np.where(idx >= 0, idx + l_p, idx + r_s)


class OracleSystem:
    def __init__(self, system):
        lo = system.lowering
        self.p = [OPhase(p) for p in system.p]
        self.l_p, self.r_p = lo.l_p, lo.r_p
        self.l_s, self.r_s, self.L = lo.l_s, lo.r_s, lo.r_s
        self.n_s = system.n_s
        self.F_o = OFunc(lo.F_o)
        self.F_c = [OFunc(f) for f in lo.F_c]
        self.n_c = len(self.F_c)
        self.which_o, self.which_c = lo.which_o, lo.which_c
        self.n_sym = lo.n_int_total + self.n_s
        self._build()

    def _build(self):  # systembase.py:366-551
        self.static = [ONode().leaf(self.l_s + k) for k in range(self.n_s)]
        self.integral = []
        for pi, p in enumerate(self.p):
            tr = lambda a: _translate(a, self.l_p[pi], self.r_s)
            for k in range(p.n_I):
                nd = ONode()
                for key, on in (("f", p.ms.f), ("m", True), ("b", p.ms.b)):
                    if on:
                        src = p.int[key][k]
                        nd.G_i += [tr(a) for a in src.G_i]
                        nd.H_i_row += [tr(a) for a in src.H_i_row]
                        nd.H_i_col += [tr(a) for a in src.H_i_col]
                self.integral.append(nd)
        basic = self.integral + self.static
        self.node_o = ONode(1, basic).bind(self.F_o)
        self.node_c = [ONode(1, basic).bind(f) for f in self.F_c]
        gradient_indices([self.node_o] + self.node_c)
        hessian_indices_system([self.node_o] + self.node_c)
        self.grad_col = _cat(self.node_o.G_i, np.int64)
        self.hess_o_row = _cat(self.node_o.H_i_row, np.int64)
        self.hess_o_col = _cat(self.node_o.H_i_col, np.int64)
        jr = [np.full(len(g), i) for i, nd in enumerate(self.node_c) for g in nd.G_i]
        jc = [g for nd in self.node_c for g in nd.G_i]
        hr = [a for nd in self.node_c for a in nd.H_i_row]
        hc = [a for nd in self.node_c for a in nd.H_i_col]
        c_ = self.n_c
        for pi, p in enumerate(self.p):
            tr = lambda a: _translate(a, self.l_p[pi], self.r_s)
            jr.append(c_ + p.jac_dyn_row)
            jc.append(tr(p.jac_dyn_col))
            c_ += p.col.n_rows * p.n_x
            jr.append(c_ + p.jac_pc_row)
            jc.append(tr(p.jac_pc_col))
            c_ += p.n_c * p.L_m
            hr += [tr(p.hess_dyn_row), tr(p.hess_pc_row)]
            hc += [tr(p.hess_dyn_col), tr(p.hess_pc_col)]
        self.m = c_
        self.jac_row, self.jac_col = _cat(jr, np.int64), _cat(jc, np.int64)
        self.hess_c_row, self.hess_c_col = _cat(hr, np.int64), _cat(hc, np.int64)

    # -- structures
    def jacobianstructure(self):
        return self.jac_row, self.jac_col

    def hessianstructure_o(self):
        return self.hess_o_row, self.hess_o_col

    def hessianstructure_c(self):
        return self.hess_c_row, self.hess_c_col

    def hessianstructure(self):
        return (
            np.concatenate([self.hess_o_row, self.hess_c_row]),
            np.concatenate([self.hess_o_col, self.hess_c_col]),
        )

    # -- values
    def _split(self, x):
        s = x[self.l_s : self.r_s]
        return s, [x[self.l_p[i] : self.r_p[i]] for i in range(len(self.p))]

    def _value_basic(self, which, x):  # :592-600
        s, xs = self._split(x)
        v = np.empty(self.n_sym)
        k = 0
        for pi, p in enumerate(self.p):
            v[k : k + p.n_I] = p.value_integral(which[pi], xs[pi], s)
            k += p.n_I
        v[k:] = s
        return v

    def error_estimation_data(self, x):
        """Per phase ``(T_x_aug, I_f_aug)`` (``phasebase.py:1355-1366``)."""
        s, xs = self._split(x)
        return [p.error_estimation_data_continuous(x_, s) for p, x_ in zip(self.p, xs)]

    def objective(self, x):  # :602-605
        return self.F_o.F(self._value_basic(self.which_o, x), 1)[0]

    def constraints(self, x):  # :607-623
        vb = self._value_basic(self.which_c, x)
        s, xs = self._split(x)
        out = [np.array([f.F(vb, 1)[0] for f in self.F_c])]
        for p, x_ in zip(self.p, xs):
            out += [p.value_dynamic(x_, s), p.value_path(x_, s)]
        return _cat(out)

    def _lists(self, which, x, second):  # :625-644 / :695-724
        s, xs = self._split(x)
        n_ = 0
        for pi, p in enumerate(self.p):
            p.integral_lists(which[pi], xs[pi], s, second)
            w, ms = p.col.w_m, p.ms
            for k in range(p.n_I):
                if which[pi][k]:
                    for attr in ("G", "H") if second else ("G",):
                        vals = []
                        if ms.f:
                            vals += [v * w[0] for v in getattr(p.int["f"][k], attr)]
                        vals += [v * w[ms.l_m : ms.r_m] for v in getattr(p.int["m"][k], attr)]
                        if ms.b:
                            vals += [v * w[-1] for v in getattr(p.int["b"][k], attr)]
                        setattr(self.integral[n_ + k], attr, vals)
            n_ += p.n_I

    def gradient(self, x):  # :646-657
        self._lists(self.which_o, x, False)
        vb = self._value_basic(self.which_o, x)
        self.node_o.g = self.F_o.G(vb, 1)
        gradient_values([self.node_o])
        grad = np.zeros(self.L)
        for i, v in zip(self.node_o.G_i, self.node_o.G):
            np.add.at(grad, i, v)
        return grad

    def jacobian(self, x):  # :659-693
        self._lists(self.which_c, x, False)
        vb = self._value_basic(self.which_c, x)
        for nd, f in zip(self.node_c, self.F_c):
            nd.g = f.G(vb, 1)
        gradient_values(self.node_c)
        out = [g for nd in self.node_c for g in nd.G]
        s, xs = self._split(x)
        for p, x_ in zip(self.p, xs):
            out += [p.grad_dynamic(x_, s), p.grad_path(x_, s)]
        return _cat(out)

    def hessian_o(self, x):  # :735-756
        self._lists(self.which_o, x, True)
        vb = self._value_basic(self.which_o, x)
        self.node_o.g, self.node_o.h = self.F_o.G(vb, 1), self.F_o.H(vb, 1)
        gradient_values([self.node_o])
        hessian_values_system([self.node_o])
        return _cat(self.node_o.H)

    def hessian_c(self, x, lam):  # :767-809
        self._lists(self.which_c, x, True)
        vb = self._value_basic(self.which_c, x)
        for nd, f in zip(self.node_c, self.F_c):
            nd.g, nd.h = f.G(vb, 1), f.H(vb, 1)
        gradient_values(self.node_c)
        hessian_values_system(self.node_c)
        out = [_cat(nd.H) * lam[i] for i, nd in enumerate(self.node_c)]
        s, xs = self._split(x)
        f_ = self.n_c
        for p, x_ in zip(self.p, xs):
            nd = p.col.n_rows * p.n_x
            out.append(p.hess_dynamic(x_, s, lam[f_ : f_ + nd]))
            f_ += nd
            out.append(p.hess_path(x_, s, lam[f_ : f_ + p.n_c * p.L_m]))
            f_ += p.n_c * p.L_m
        return _cat(out)

    def hessian(self, x, lam, sigma):  # :820-835
        return np.concatenate([self.hessian_o(x) * sigma, self.hessian_c(x, lam)])
