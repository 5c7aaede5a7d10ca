"""Collocation tables for one phase: nodes, weights, variable layout and the
block-structured translation / integration operators.

This is the *data* side of the hot path.  The reference builds
the same tables in ``pockit/lobatto/discretization.py:80-227, 414-441`` (LGL)
and ``pockit/radau/discretization.py:89-257, 488-521`` (LGR) on top of
``pockit/base/discretizationbase.py:98-329``.  Here everything is kept in the
form the device engine consumes: per-interval dense blocks plus flat COO
triplets already split into front / middle / back column classes.

Numerical note (measured): the reference obtains the LGL / LGR
abscissae from ``numpy.roots`` of monomial-basis polynomials, which loses
~1e-10 at 20 points.  Value parity at 1e-12 therefore requires the *same*
root-finding route, so `gauss_lobatto` / `gauss_radau` deliberately use it too.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass

import numpy as np
import scipy.special

__all__ = ["Collocation", "gauss_lobatto", "gauss_radau", "antiderivative_block"]


# ----------------------------------------------------------------------------
# abscissae / weights on [-1, 1]
# ----------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def gauss_lobatto(n: int) -> tuple[np.ndarray, np.ndarray]:
    """LGL points/weights (reference: ``lobatto/discretization.py:80-110``)."""
    if n <= 0:
        raise ValueError("Number of interpolation points must be at least 1.")
    if n == 1:
        return np.array([0.0]), np.array([2.0])
    deg = n - 1
    legendre = scipy.special.legendre(deg)
    interior = [r.real for r in np.roots(np.polyder(legendre))]
    pts = np.array(sorted([-1.0, *interior, 1.0]), dtype=np.float64)
    edge = 2.0 / deg / (deg + 1)
    wts = np.empty(n, dtype=np.float64)
    wts[0] = wts[-1] = edge
    for k in range(1, n - 1):
        wts[k] = 2.0 / deg / (deg + 1) / np.polyval(legendre, pts[k]) ** 2
    pts.setflags(write=False)
    wts.setflags(write=False)
    return pts, wts


@functools.lru_cache(maxsize=None)
def gauss_radau(n: int) -> tuple[np.ndarray, np.ndarray]:
    """LGR points/weights, left end included (``radau/discretization.py:89-114``)."""
    if n <= 0:
        raise ValueError("Number of interpolation points must be at least 1.")
    jac = scipy.special.jacobi(n - 1, 0, 1)
    pts = np.array(sorted([-1.0, *[r.real for r in np.roots(jac)]]), dtype=float)
    leg = scipy.special.legendre(n)
    wts = (1 - pts) / (n * np.polyval(leg, pts)) ** 2
    pts.setflags(write=False)
    wts.setflags(write=False)
    return pts, wts


def _lagrange_basis(at: np.ndarray, nodes: np.ndarray, bary: np.ndarray) -> np.ndarray:
    """Barycentric Lagrange basis ``L_j(at_k)`` (``discretizationbase.py:41-95``)."""
    if len(nodes) == 1:
        return np.ones((len(at), 1))
    gap = at[:, None] - nodes[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        terms = bary[None, :] / gap
        denom = terms.sum(axis=1)
        basis = terms / denom[:, None]
    basis[np.isclose(denom, 0.0) | ~np.isfinite(denom), :] = 0.0
    hit_k, hit_j = np.nonzero(np.isclose(at[:, None], nodes[None, :], rtol=1e-13, atol=1e-13))
    seen = set()
    for k, j in zip(hit_k, hit_j):  # first coincident node wins, like the reference
        if k in seen:
            continue
        seen.add(k)
        basis[k, :] = 0.0
        basis[k, j] = 1.0
    return basis


def antiderivative_block(nodes_in: np.ndarray, nodes_out: np.ndarray) -> np.ndarray:
    """``B[r, j] = ∫_1^{nodes_out[r]} L_j(τ) dτ`` for the Lagrange basis on ``nodes_in``.

    Same construction as ``discretizationbase.py:98-180``: barycentric weights,
    Gauss–Legendre rule with ``max(30, 3n)`` points mapped onto ``[1, x_r]``.
    """
    nodes_in = np.asarray(nodes_in, dtype=np.float64)
    nodes_out = np.asarray(nodes_out, dtype=np.float64)
    n, m = len(nodes_in), len(nodes_out)
    out = np.zeros((m, n))
    if n == 0 or m == 0:
        return out
    bary = np.ones(n)
    for j in range(n):
        for k in range(n):
            if k != j:
                bary[j] /= nodes_in[j] - nodes_in[k]
    gx, gw = np.polynomial.legendre.leggauss(max(30, 3 * n))
    for r in range(m):
        target = nodes_out[r]
        if np.isclose(target, 1.0, rtol=1e-13, atol=1e-13):
            continue
        half = 0.5 * (target - 1.0)
        centre = 0.5 * (target + 1.0)
        out[r, :] = np.dot(half * gw, _lagrange_basis(half * gx + centre, nodes_in, bary))
    return out


@functools.lru_cache(maxsize=None)
def _unit_integration_block(scheme: str, n: int) -> np.ndarray:
    if scheme == "lgl":
        pts, _ = gauss_lobatto(n)
        blk = antiderivative_block(pts, pts[:-1])  # (n-1) x n   lobatto:155-166
    else:
        pts, _ = gauss_radau(n)
        blk = antiderivative_block(pts, pts)  # n x n             radau:185-196
    blk.setflags(write=False)
    return blk


# ----------------------------------------------------------------------------
# flat operators
# ----------------------------------------------------------------------------
@dataclass
class Triplets:
    """COO triplets of one column class (front / middle / back)."""

    row: np.ndarray  # int32
    col: np.ndarray  # int32
    data: np.ndarray  # float64
    k: np.ndarray  # int64: position of each triplet in the un-split operator

    def __len__(self) -> int:
        return len(self.row)


@dataclass
class SplitOperator:
    """An operator split by column class like ``CooMatrixNode``
    (``discretizationbase.py:258-329``): ``f`` hits the front node, ``b`` the back
    node, ``m`` everything else.  Order inside each part is row-major."""

    f: Triplets
    m: Triplets
    b: Triplets
    shape: tuple[int, int]

    def dot(self, v: np.ndarray) -> np.ndarray:
        """Row-sequential product (same accumulation order as SciPy CSR)."""
        out = np.zeros(self.shape[0])
        rows = np.concatenate([self.f.row, self.m.row, self.b.row])
        cols = np.concatenate([self.f.col, self.m.col, self.b.col])
        data = np.concatenate([self.f.data, self.m.data, self.b.data])
        order = np.concatenate([self.f.k, self.m.k, self.b.k]).argsort(kind="stable")
        np.add.at(out, rows[order], data[order] * v[cols[order]])
        return out


@dataclass
class NodeClass:
    """front index / middle range / back index of a slab (``IndexNode``,
    ``discretizationbase.py:199-255``)."""

    front: int | None
    lo: int
    hi: int
    back: int | None

    @property
    def f(self) -> bool:
        return self.front is not None

    @property
    def b(self) -> bool:
        return self.back is not None

    @property
    def m(self) -> slice:
        return slice(self.lo, self.hi)

    @property
    def l_m(self) -> int:
        return self.lo

    @property
    def r_m(self) -> int:
        return self.hi

    @property
    def L_m(self) -> int:
        return self.hi - self.lo


def _split(row, col, data, nodes: NodeClass, shape) -> SplitOperator:
    keep = data != 0.0
    row, col, data = row[keep], col[keep], data[keep]
    k = np.arange(len(row), dtype=np.int64)
    front = col == (nodes.front if nodes.f else -1)
    back = (col == (nodes.back if nodes.b else -1)) & ~front
    mid = ~(front | back)

    def part(mask):
        return Triplets(
            row[mask].astype(np.int32), col[mask].astype(np.int32), data[mask].astype(np.float64), k[mask]
        )

    return SplitOperator(part(front), part(mid), part(back), shape)


class Collocation:
    """All per-phase tables for a mesh (``mesh`` scaled to [0, 1]) with
    ``num_point[k]`` collocation points in interval ``k``.

    Layout: a scalar state owns ``L_x`` consecutive slots, a
    control ``L_m``; LGL shares interval borders (``L_x == L_m``), LGR appends
    the terminal node to every state (``L_x == L_m + 1``).
    """

    def __init__(self, scheme: str, mesh: np.ndarray, num_point: np.ndarray, n_x: int, n_u: int):
        if scheme not in ("lgl", "lgr"):
            raise ValueError("scheme must be 'lgl' or 'lgr'")
        self.scheme = scheme
        self.mesh = np.asarray(mesh, dtype=np.float64)
        self.num_point = np.asarray(num_point, dtype=np.int32)
        self.n_x, self.n_u = int(n_x), int(n_u)
        npt = self.num_point.astype(np.int64)
        n_int = len(npt)
        width = np.diff(self.mesh)
        centre = (self.mesh[1:] + self.mesh[:-1]) / 2
        self.width = width

        if scheme == "lgl":
            # interval k covers nodes [l_m[k], l_m[k] + n_k), sharing its ends
            l_m = np.concatenate(([0], np.cumsum(npt[:-1] - 1)))
            rows_per = npt - 1
            rule = gauss_lobatto
        else:
            l_m = np.concatenate(([0], np.cumsum(npt[:-1])))
            rows_per = npt
            rule = gauss_radau
        r_m = l_m + npt
        self.l_m, self.r_m = l_m, r_m
        self.L_m = int(r_m[-1])
        self.L_x = self.L_m if scheme == "lgl" else self.L_m + 1
        self.row_start = np.concatenate(([0], np.cumsum(rows_per)))  # defect rows / state
        self.n_rows = int(self.row_start[-1])

        t_m = np.zeros(self.L_m)
        w_m = np.zeros(self.L_m)
        for k in range(n_int):
            pts, wts = rule(int(npt[k]))
            t_m[l_m[k] : r_m[k]] = pts * width[k] / 2 + centre[k]
            if scheme == "lgl":
                w_m[l_m[k] : r_m[k]] += wts * width[k] / 2  # shared borders add up
            else:
                w_m[l_m[k] : r_m[k]] = wts * width[k] / 2
        self.t_m, self.w_m = t_m, w_m

        # variable slabs
        sizes = [self.L_x] * self.n_x + [self.L_m] * self.n_u
        ends = np.cumsum(sizes) if sizes else np.array([], dtype=np.int64)
        self.r_v = np.asarray(ends, dtype=np.int64)
        self.l_v = self.r_v - np.asarray(sizes, dtype=np.int64)
        self.L_xu = int(ends[-1]) if len(ends) else 0
        self.L = self.L_xu + 2
        self.l_d = np.arange(self.n_x, dtype=np.int64) * self.n_rows
        self.r_d = self.l_d + self.n_rows

        if scheme == "lgl":
            self.index_state = NodeClass(0, 1, self.L_m - 1, self.L_m - 1)
            self.index_control = NodeClass(0, 1, self.L_m - 1, self.L_m - 1)
            self.index_mstage = NodeClass(0, 1, self.L_m - 1, self.L_m - 1)
        else:
            self.index_state = NodeClass(0, 1, self.L_m, self.L_m)
            self.index_control = NodeClass(0, 1, self.L_m, None)
            self.index_mstage = NodeClass(0, 1, self.L_m, None)

        # operators: T = [I | -1] per interval, I = antiderivative * width / 2
        t_row, t_col, t_val = [], [], []
        i_row, i_col, i_val = [], [], []
        self.blocks: list[np.ndarray] = []  # scaled integration block of every interval
        for k in range(n_int):
            n = int(npt[k])
            rows = int(rows_per[k])
            r0 = int(self.row_start[k])
            c0 = int(l_m[k])
            # translation block: rows x (rows + 1)
            eye_r = r0 + np.arange(rows)
            t_row.append(np.repeat(eye_r, 2))
            t_col.append(np.stack([c0 + np.arange(rows), np.full(rows, c0 + rows)], axis=1).ravel())
            t_val.append(np.tile([1.0, -1.0], rows))
            blk = _unit_integration_block(scheme, n) * width[k] / 2
            self.blocks.append(blk)
            i_row.append(np.repeat(eye_r, n))
            i_col.append(np.tile(c0 + np.arange(n), rows))
            i_val.append(blk.ravel())
        t_row, t_col, t_val = map(np.concatenate, (t_row, t_col, t_val))
        i_row, i_col, i_val = map(np.concatenate, (i_row, i_col, i_val))
        # [1 ... -1] rows must be column-sorted; for rows == 1 in LGL the two
        # columns are already ascending, so the generic layout is row-major.
        self.T = _split(t_row, t_col, t_val, self.index_state, (self.n_rows, self.L_x))
        self.I = _split(i_row, i_col, i_val, self.index_mstage, (self.n_rows, self.L_m))
        # fast-path predicate for the engine: every interval has the same order and
        # no entry of any block was dropped as an exact zero (widths may differ;
        # the reference scales each block by its own width, (unit * d) / 2)
        self.same_order = bool(np.all(npt == npt[0]))
        self.dense_blocks = bool(np.count_nonzero(i_val) == len(i_val))

    # node -> interval map, useful to the engine for non-uniform meshes
    def interval_of_row(self) -> np.ndarray:
        return np.repeat(np.arange(len(self.num_point)), np.diff(self.row_start))


# ----------------------------------------------------------------------------
# augmented mesh (one more point per interval): data of the continuous error estimate
# ----------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _lagrange_values(nodes: tuple, at: tuple) -> np.ndarray:
    """``V[k, i]`` = i-th Lagrange basis polynomial on ``nodes`` evaluated at ``at[k]``.

    The reference builds these tables with ``scipy.interpolate.lagrange`` (monomial basis) and
    ``numpy.polyval`` (``lobatto/discretization.py:256-301``, ``radau/discretization.py:286-355``);
    that route is ill-conditioned at 20 points, so parity requires the same library calls."""
    import warnings

    import scipy.interpolate

    nodes_a, at_a = np.array(nodes, dtype=np.float64), np.array(at, dtype=np.float64)
    cols = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)  # SciPy deprecates `lagrange`; the reference uses it
        for i in range(len(nodes_a)):
            y = np.zeros(len(nodes_a))
            y[i] = 1
            cols.append(np.polyval(scipy.interpolate.lagrange(nodes_a, y), at_a))
    return np.array(cols, dtype=np.float64).T


class AugmentedCollocation:
    """Operators of the continuous error estimate (``phasebase.py:1339-1366``) as CSR triplets:

    * ``V``  interpolation of every state / control from the mesh to the augmented mesh
      (``n_k + 1`` points per interval), block diagonal over the variables (``V_xu_aug``);
    * ``T``  the same interpolant minus its value at the interval end (``T_x_aug``), states only;
    * ``I``  the integration operator of the augmented mesh (``I_m(mesh, num_point + 1)``);
    * ``t_m`` / ``l_m`` / ``r_m`` of the augmented mesh.

    Every operator is applied as a CSR matrix-vector product with sequential row sums, exactly
    like ``scipy.sparse.csr_array.dot`` in the reference."""

    def __init__(self, col: Collocation):
        import scipy.sparse as sp

        self.col = col
        scheme, mesh, npt = col.scheme, col.mesh, col.num_point.astype(np.int64)
        aug = Collocation(scheme, mesh, npt + 1, col.n_x, col.n_u)
        self.aug = aug
        self.t_m, self.l_m, self.r_m, self.L_m = aug.t_m, aug.l_m, aug.r_m, aug.L_m
        n_int = len(npt)
        lgl = scheme == "lgl"
        rule = gauss_lobatto if lgl else gauss_radau

        def assemble(blocks, rows0, cols0, shape):
            r, c, d = [], [], []
            for blk, r0, c0 in zip(blocks, rows0, cols0):
                nr, nc = blk.shape
                r.append(r0 + np.repeat(np.arange(nr), nc))
                c.append(c0 + np.tile(np.arange(nc), nr))
                d.append(blk.ravel())
            m = sp.coo_array((np.concatenate(d), (np.concatenate(r), np.concatenate(c))), shape=shape)
            m.sum_duplicates()
            m.eliminate_zeros()
            return m.tocsr()

        col0 = col.l_m  # first mesh node of every interval (LGL and LGR alike)
        if lgl:
            # rows: shared-border layout of the augmented mesh; the first row of every interval but
            # the first is dropped (its value comes from the previous interval's last row)
            vb, vr0, tb = [], [], []
            for k in range(n_int):
                n = int(npt[k])
                x, xa = rule(n)[0], rule(n + 1)[0]
                V = _lagrange_values(tuple(x), tuple(xa))            # (n+1) x n
                vb.append(V if k == 0 else V[1:])
                vr0.append(int(aug.l_m[k]) + (0 if k == 0 else 1))
                tb.append(V[:-1] - V[-1])                             # n x n
            Vs = assemble(vb, vr0, col0, (aug.L_m, col.L_m))
            Vx = Vu = Vs
            t_rows0 = np.concatenate(([0], np.cumsum(npt[:-1])))
            Ts = assemble(tb, t_rows0, col0, (int(npt.sum()), col.L_m))
        else:
            vxb, vub, tb = [], [], []
            for k in range(n_int):
                n = int(npt[k])
                x, xa = rule(n)[0], rule(n + 1)[0]
                x1 = tuple(np.concatenate((x, [1.0])))
                vxb.append(_lagrange_values(x1, tuple(xa)))           # (n+1) x (n+1), states carry the interval end
                vub.append(_lagrange_values(tuple(x), tuple(xa)))     # (n+1) x n
                Tv = _lagrange_values(x1, tuple(np.concatenate((xa, [1.0]))))
                tb.append(Tv[:-1] - Tv[-1])                           # (n+1) x (n+1)
            Vx = assemble(vxb, aug.l_m, col0, (aug.L_m, col.L_x))
            Vu = assemble(vub, aug.l_m, col0, (aug.L_m, col.L_m))
            Ts = assemble(tb, aug.l_m, col0, (aug.L_m, col.L_x))
        self.V = sp.block_diag([Vx] * col.n_x + [Vu] * col.n_u).tocsr() if col.n_x + col.n_u else sp.csr_array((0, 0))
        self.T = sp.block_diag([Ts] * col.n_x).tocsr() if col.n_x else sp.csr_array((0, 0))
        self.rows = Ts.shape[0]  # rows per state of T.x and of I.f
        I = aug.I
        trip = [I.f, I.m, I.b]
        Im = sp.coo_array((np.concatenate([t.data for t in trip]), (np.concatenate([t.row for t in trip]),
                                                                   np.concatenate([t.col for t in trip]))),
                          shape=(aug.n_rows, aug.L_m))
        self.I = Im.tocsr()
        self.I.sort_indices()
        assert self.I.shape[0] == self.rows
