"""Device plan: turns a :class:`~pockit_b200.system.SystemLowering` into what the
C-ABI engine consumes -- CUDA C for the per-node programs, job records for the
hand-written kernels, constant pools and the scalar / node-table layouts.

Nothing here depends on the mesh *size* except numbers stored in tables, so a
re-meshed problem re-uses the generated source verbatim (and the engine's
cubin cache): sizes, offsets and slot bases reach the kernels through a
``__constant__`` table indexed by literals.

Data layout in HBM (per engine, ``B`` instances)::

    X      [B][L]        optimisation vectors                 (input)
    LAM    [B][m]        constraint multipliers, SIG [B]      (input, Hessian only)
    OUT    [B][n_out]    objective / gradient / constraints / Jacobian values / Hessian values
    S      [B][n_scalar] scalar table: dt, substituted boundary values, front/back/basic list
                         values, reduction results, system-level derivative leaves
    W      rows of [B][L_m(phase)]  node table: one row per per-node list that is later expanded
                         through the integration operator, reduced, or scaled
    pools  double / int64 constants: integration blocks, triplets, mesh fractions, weights
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import sympy as sp

from .chain import Leaf, Term
from .emit import PRELUDE, emit_block
from .phase import BcType, Segment
from .system import SysList, SysSegment, SystemLowering

MODES = ("objective", "constraints", "gradient", "jacobian", "hessian", "set")
OBJ, CONS, GRAD, JAC, HESS, SET = range(6)
# SET = all five callbacks at one (x, lambda, sigma) as ONE pipeline: a single per-node program
# evaluates every leaf once (the Hessian's leaves are a superset of the others'), one reduction and
# one system program feed all consumers, and the outputs land in one buffer laid out
# [objective | gradient | constraints | Jacobian values | Hessian values] (SET_ORDER).
SET_ORDER = (OBJ, GRAD, CONS, JAC, HESS)
# The pipeline may also cover a subset (DevicePlan(set_subs=...)); the engine's default is the three
# latency-bound callbacks SMALL_SET, whose separate chains of 3-5 small kernels each become ONE chain
# (one per-node program with a shared CSE, one reduction, one system program, defects, gradient
# gather) while the Jacobian and the Hessian keep their own pipelines beside it.
SMALL_SET = (OBJ, GRAD, CONS)
ST_REDUCE, ST_DEFECT, ST_GENERIC, ST_EXPAND, ST_GRAD_RANGE, ST_GRAD_SCALAR = range(6)
J_CONST, J_KRON, J_EXPAND_TABLE, J_SCALED, J_SYS, J_OUTER, J_TRIL = range(7)
F_A_SCALAR, F_B_SCALAR, F_A_UNIT, F_B_UNIT, F_LAM = 1, 2, 4, 8, 16

JOB_DTYPE = np.dtype([("type", "<i4"), ("flags", "<i4"), ("i", "<i8", (16,)), ("f", "<f8", (2,))])

NODE_BLOCK = 128
SPLIT_MIN = 1024  # slot runs shorter than this are not split between mesh shards (rank 0 owns them)


class Pools:
    """Append-only constant pools with de-duplication of identical arrays."""

    def __init__(self):
        self.d: list[np.ndarray] = []
        self.i: list[np.ndarray] = []
        self.nd = 0
        self.ni = 0
        self._seen = {}

    def _add(self, arr, kind):
        arr = np.ascontiguousarray(arr, dtype=np.float64 if kind == "d" else np.int64)
        key = (kind, arr.shape, arr.tobytes())
        if key in self._seen:
            return self._seen[key]
        if kind == "d":
            off = self.nd
            self.d.append(arr)
            self.nd += arr.size
        else:
            off = self.ni
            self.i.append(arr)
            self.ni += arr.size
        self._seen[key] = off
        return off

    def dbl(self, arr) -> int:
        return self._add(arr, "d")

    def int(self, arr) -> int:
        return self._add(arr, "i")

    def arrays(self):
        d = np.concatenate(self.d) if self.d else np.zeros(1)
        i = np.concatenate(self.i) if self.i else np.zeros(1, dtype=np.int64)
        return np.ascontiguousarray(d), np.ascontiguousarray(i)


_FN_LEAF = re.compile(r"\b[FGH]_([dIc])_(\d+)(?:_\d+)?\b")


def _function_of(code) -> Optional[tuple]:
    """(family, index) of the function whose leaves a store term uses (each term uses one), or None."""
    if not isinstance(code, str):
        return None
    m = _FN_LEAF.search(code)
    return (m.group(1), int(m.group(2))) if m else None


def _ident(leaf: Leaf) -> str:
    return leaf.kind + "_" + "_".join(str(k) for k in leaf.key) if leaf.key else leaf.kind


def term_code(t: Term, weight: bool = False) -> str:
    """C expression of a product tree, multiplied in the reference's association."""

    def walk(tree):
        if isinstance(tree, Leaf):
            return _ident(tree)
        return f"({walk(tree.left)} * {walk(tree.right)})"

    body = "1.0" if t.tree is None else walk(t.tree)
    if t.coef == -1.0:
        body = f"(-{body})"
    elif t.coef != 1.0:
        body = f"({body} * {t.coef!r})"
    return f"({body} * wq)" if weight else body


@dataclass
class PhaseProgram:
    """Everything one phase's per-node kernel has to compute in one mode."""

    need: set = field(default_factory=set)  # function leaves (kind, fam, i, k)
    s_store: dict = field(default_factory=dict)  # (nset, code) -> scalar slot
    w_store: dict = field(default_factory=dict)  # (nset, code) -> table row id
    direct: list = field(default_factory=list)  # (nset, code, dst, lam or -1, c_lo)
    walks: list = field(default_factory=list)  # (code, dst, lam row base or -1, sign): fused block expansion


class ModePlan:
    """Plan of one callback."""

    def __init__(self, owner: "DevicePlan", mode: int):
        self.owner, self.mode = owner, mode
        lo = owner.lo
        self.prog = [PhaseProgram() for _ in lo.phases]
        self.sys_need: dict = {}  # sys Leaf -> scalar slot
        self.jobs = {s: [] for s in range(6)}
        self.rows: list[int] = []  # phase of every table row
        self.n_scalar = owner.header_slots
        self.table: list[int] = []  # __constant__ table entries
        self.int_needed = [np.zeros(p.n_I, dtype=bool) for p in lo.phases]
        self.n_out = 0
        self.src = None
        self.expand_groups: dict = {}
        self.owned_runs: Optional[list] = None  # mesh shard: (offset, count) runs of the output computed here
        self.node_threads = [1] * len(lo.phases)  # threads per node of every phase's program (expression groups)
        self.sub = mode   # callback currently being planned (differs from `mode` only inside SET)
        self.base = 0     # first output slot of that callback
        self.sub_range: dict = {}  # SET: callback -> (offset, count) inside the combined output
        self.grad_range = (0, 0)   # output slots the gradient gather-sums into (zeroed first)
        self._build()
        if owner.shard is not None:
            self._shard(owner.shard[0], owner.shard[1])

    # ---------------------------------------------------------------- allocation helpers
    def tab(self, value) -> int:
        if isinstance(value, tuple):  # ('row', r): base of a node-table row
            value = self.row_base(value[1])
        self.table.append(int(value))
        return len(self.table) - 1

    def row_base(self, r: int) -> int:
        lo, B = self.owner.lo, self.owner.B
        return sum(B * lo.phases[ph].L_m for ph in self.rows[:r])

    @property
    def n_table(self) -> int:
        return self.row_base(len(self.rows))

    @property
    def n_scalar_final(self) -> int:
        return self.n_scalar

    def new_row(self, phase: int) -> int:
        self.rows.append(phase)
        return len(self.rows) - 1

    def new_slot(self, n: int = 1) -> int:
        s = self.n_scalar
        self.n_scalar += n
        return s

    def _track(self, pi: int, t: Term):
        for lf in t.leaves():
            if lf.kind in ("F", "G", "H"):
                self.prog[pi].need.add(lf)

    def scalar_of(self, pi: int, nset: str, t: Term, weight=False) -> int:
        """Scalar-table slot holding ``t`` evaluated at the front / back node (or a
        function of ``s`` only for the 'basic' set)."""
        self._track(pi, t)
        key = (nset, term_code(t, weight))
        store = self.prog[pi].s_store
        if key not in store:
            store[key] = self.new_slot()
        return store[key]

    def row_of(self, pi: int, nset: str, t: Term, weight=False) -> int:
        self._track(pi, t)
        key = (nset, term_code(t, weight))
        store = self.prog[pi].w_store
        if key not in store:
            store[key] = self.new_row(pi)
        return store[key]

    def sys_slot(self, t: Term) -> int:
        (lf,) = t.leaves()
        if lf not in self.sys_need:
            self.sys_need[lf] = self.new_slot()
        return self.sys_need[lf]

    def job(self, stage: int, typ: int = 0, flags: int = 0, f0: float = 1.0, **iv) -> dict:
        rec = {"type": typ, "flags": flags, "i": [0] * 16, "f": [f0, 0.0], "rows": {}}
        for k, v in iv.items():
            idx = int(k[1:])
            if isinstance(v, tuple) and v[0] == "row":  # table-row base, resolved once B is known
                rec["rows"][idx] = v[1]
            else:
                rec["i"][idx] = int(v)
        self.jobs[stage].append(rec)
        return rec

    # ---------------------------------------------------------------- mesh sharding
    def _shard(self, g: int, G: int):
        """Keep only rank ``g``'s share of the slot-streaming work ("one very fine mesh").

        The cheap, latency-bound part -- per-node programs, quadrature sums, system program -- is
        *replicated* on every rank (a few microseconds on 40 k threads, and it makes the integrals
        bit-identical everywhere without an all-reduce); the HBM- and PCIe-bound part is split:
        every block expansion keeps a contiguous range of intervals, long table / constant /
        scaled runs keep a contiguous sub-range, short runs and the small callbacks
        (objective, gradient, constraints) stay on rank 0.  ``owned_runs`` lists the output
        slots this rank produces; over all ranks they partition ``[0, n_out)``."""
        lo = self.owner.lo
        if self.mode not in (JAC, HESS):
            self.owned_runs = [(0, self.n_out)] if g == 0 else []
            if g != 0:
                self.jobs = {s: [] for s in range(6)}
            return

        # share boundaries: equal shares by default; with `shard = (rank, world, weights)` rank r gets a share
        # proportional to weights[r] (the ranks' device-to-host links are not equally fast on every box, and
        # the end-to-end path is copy-bound: meshshard measures the rates and passes them in).  Every rank
        # derives the same boundaries from the same vector, so the shares still partition the output.
        cum = self.owner.shard_cum

        def cut(count, r):
            return count if r >= G else (0 if r <= 0 else min(count, int(count * cum[r])))

        def part(count):
            return cut(count, g), cut(count, g + 1)

        def clone(rec):
            r = dict(rec)
            r["i"] = list(rec["i"])
            r["f"] = list(rec["f"])
            r["rows"] = dict(rec["rows"])
            return r

        runs = []
        generic = []
        for rec in self.jobs[ST_GENERIC]:
            t, dst, cnt = rec["type"], rec["i"][0], rec["i"][1]
            splittable = cnt >= SPLIT_MIN and (
                t in (J_CONST, J_EXPAND_TABLE) or (t == J_SCALED and not rec["flags"] & F_A_SCALAR)
            )
            if not splittable:
                if g == 0:
                    generic.append(rec)
                    runs.append((dst, cnt))
                continue
            e0, e1 = part(cnt)
            if e1 == e0:
                continue
            r = clone(rec)
            r["i"][0], r["i"][1] = dst + e0, e1 - e0
            if t == J_CONST:
                r["i"][6] += e0
            elif t == J_EXPAND_TABLE:
                for k in (6, 7, 8):
                    r["i"][k] += e0
            else:
                r["i"][8] += e0
            generic.append(r)
            runs.append((dst + e0, e1 - e0))
        self.jobs[ST_GENERIC] = generic
        # Block expansions: the (list, interval) tiles of all jobs, in output order (list-major), are
        # dealt out as ONE contiguous range per rank, weighted by tile size.  A rank then owns a few
        # whole lists plus at most two partial ones -- long contiguous runs, i.e. few large
        # device-to-host copies (splitting every list by interval range instead gave ~70 copies of
        # 1.6 MB per rank and set at 4 GPUs).
        tiles = []  # (job record, list index, intervals, slots per tile)
        for rec in self.jobs[ST_EXPAND]:
            n, rows = rec["i"][3], rec["i"][4]
            for li in range(len(rec["lists"])):
                tiles.append((rec, li, rec["i"][11] // n, n * rows))
        total = sum(nK * sz for _, _, nK, sz in tiles)
        lo_w, hi_w = cut(total, g), cut(total, g + 1)
        expand, seen = [], 0
        groups: dict = {}  # (id(job record), Ka, Kb) -> cloned record holding the lists with that range
        for rec, li, nK, sz in tiles:
            first, last = seen, seen + nK * sz
            seen = last
            a, b = max(first, lo_w), min(last, hi_w)
            if b <= a:
                continue
            # whole tiles only: a tile belongs to the rank its first slot falls to
            Ka, Kb = -(-(a - first) // sz), -(-(b - first) // sz)
            if Kb <= Ka:
                continue
            n, rows = rec["i"][3], rec["i"][4]
            key = (id(rec), Ka, Kb)
            r = groups.get(key)
            if r is None:
                r = clone(rec)
                r["i"][11] = (Kb - Ka) * n
                r["i"][6] += Ka * rec["i"][5]
                r["i"][8] += Ka
                if rec["i"][2] >= 0:
                    r["i"][2] += Ka * rows
                r["lists"] = []
                groups[key] = r
                expand.append(r)
            dst, row = rec["lists"][li]
            r["lists"].append((dst + Ka * n * rows, row))
            runs.append((dst + Ka * n * rows, (Kb - Ka) * n * rows))
        self.jobs[ST_EXPAND] = expand
        # runs the (replicated) per-node programs write themselves: split ownership of the copy only
        for pi, prog in enumerate(self.prog):
            col = lo.phases[pi].col
            for nset, _code, dst, _lam, _c_lo in prog.direct:
                cnt = {"all": col.L_m, "mid": lo.low[pi].n_mid}.get(nset, 1)
                if cnt >= SPLIT_MIN:
                    e0, e1 = part(cnt)
                    if e1 > e0:
                        runs.append((dst + e0, e1 - e0))
                elif g == 0:
                    runs.append((dst, cnt))
        runs.sort()
        merged = []
        for off, cnt in runs:
            if cnt <= 0:
                continue
            if merged and merged[-1][0] + merged[-1][1] == off:
                merged[-1] = (merged[-1][0], merged[-1][1] + cnt)
            else:
                merged.append((off, cnt))
        self.owned_runs = merged

    # ---------------------------------------------------------------- build per mode
    def _build(self):
        if self.mode == SET:
            if self.owner.shard is not None or self.owner.fused:
                raise ValueError("the set pipeline is not planned for mesh shards / the fused variant")
            off = 0
            for sub in self.owner.set_subs:
                self.sub, self.base = sub, off
                count = self._build_one(sub)
                self.sub_range[sub] = (off, count)
                off += count
            self.n_out = off
        else:
            self.n_out = self._build_one(self.mode)
            self.sub_range[self.mode] = (0, self.n_out)
        self._integral_rows()

    def _build_one(self, mode: int) -> int:
        """Plan callback ``mode`` with its outputs starting at slot ``self.base``; returns their count."""
        lo = self.owner.lo
        if mode == OBJ:
            self._need_system_value(lo.F_o, "o", lo.which_o)
            return 1
        if mode == CONS:
            for i, fn in enumerate(lo.F_c):
                self.sys_need[Leaf("sF", ("c", i))] = -1 - (self.base + i)  # written straight to the output
            self._need_integrals(lo.which_c)
            self._constraints()
            return lo.m
        if mode == GRAD:
            self._need_integrals(lo.which_o)
            self._gradient()
            self.grad_range = (self.base, lo.r_s)
            return lo.r_s
        if mode == JAC:
            self._need_integrals(lo.which_c)
            self._slots(lo.jac_segments)
            return lo.nnz_jac
        self._need_integrals([a | b for a, b in zip(lo.which_o, lo.which_c)])
        self._slots(lo.hess_o_segments + lo.hess_c_segments)
        return lo.nnz_hess_o + lo.nnz_hess_c

    def _need_system_value(self, fn, tag, which):
        self.sys_need[Leaf("sF", (tag,))] = -1 - self.base
        self._need_integrals(which)

    def _need_integrals(self, which):
        for pi, w in enumerate(which):
            self.int_needed[pi] |= w

    def _integral_rows(self):
        """Quadrature: ``I_k = dt * sum_c w_c g_k(c)`` (phasebase.py:997-1006).  The node
        program stores ``g_k(c) * w_c`` in a table row, a REDUCE job sums it."""
        lo = self.owner.lo
        self.int_sum_slot = {}
        # only integrals that a needed system-level expression really reads
        used = set()
        for lf in self.sys_need:
            fn = lo.F_o if lf.key[0] == "o" else lo.F_c[lf.key[1]]
            e = fn.expr if lf.kind == "sF" else (fn.G_expr if lf.kind == "sG" else fn.H_expr)[lf.key[-1]]
            used |= e.free_symbols
        for pi, p in enumerate(lo.phases):
            self.int_needed[pi] &= np.array([sym in used for sym in p.I], dtype=bool)
        for pi, p in enumerate(lo.phases):
            for k in range(p.n_I):
                if self.int_needed[pi][k]:
                    t = Term(Leaf("F", ("I", k)))
                    self._track(pi, t)
                    row = self.new_row(pi)
                    self.prog[pi].w_store[("all", term_code(t, True))] = row
                    slot = self.new_slot()
                    self.int_sum_slot[(pi, k)] = slot
                    self.job(ST_REDUCE, i0=("row", row), i1=p.L_m, i2=0, i3=p.L_m, i4=slot)

    # -- constraints: defects + path values
    def _constraints(self):
        lo, own = self.owner.lo, self.owner
        for pi, p in enumerate(lo.phases):
            col = p.col
            base = self.base + lo.con_base[pi]
            fd_rows = []
            for i in range(p.n_x):
                t = Term(Leaf("F", ("d", i)))
                self._track(pi, t)
                row = self.new_row(pi)
                self.prog[pi].w_store[("all", term_code(t))] = row
                fd_rows.append(row)
            if p.n_x:
                d = own.defect_tables[pi]
                self.job(
                    ST_DEFECT, 0, d["rows_blk"], i0=lo.l_p[pi], i1=col.L_x, i2=col.L_m, i3=p.n_x, i4=col.n_rows,
                    i5=d["row_ptr"], i6=d["col"], i7=d["data"], i8=d["tpos"], i9=d["tneg"],
                    i10=("row", fd_rows[0]), i11=own.header[pi], i12=base,
                    i13=d["n"], i14=d["unit"], i15=d["width"],
                )
            pbase = base + col.n_rows * p.n_x
            for q in range(p.n_c):
                t = Term(Leaf("F", ("c", q)))
                self._track(pi, t)
                self.prog[pi].direct.append(("all", term_code(t), pbase + q * col.L_m, -1, 0))

    # -- gradient: gather-sum of the objective's lists (systembase.py:646-657)
    def _gradient(self):
        lo = self.owner.lo
        ranges: dict = {}
        scalars: dict = {}
        for sg in lo.grad_segments:
            if sg.count == 0:
                continue
            g = self.sys_slot(sg.sys)
            if sg.kind == "sys":
                scalars.setdefault(int(sg.cols[0]), []).append((-1, g))
                continue
            pi, p = sg.phase, lo.phases[sg.phase]
            if sg.nset != "mid":
                scalars.setdefault(int(sg.cols[0]), []).append((self.scalar_of(pi, sg.nset, sg.term, True), g))
                continue
            row = self.row_of(pi, "mid", sg.term, True)
            c_lo = lo.low[pi].mid_lo
            if sg.count == 1 or sg.cols[0] == sg.cols[-1]:  # one column: np.add.at accumulates into it
                slot = self.new_slot()
                self.job(ST_REDUCE, i0=("row", row), i1=p.L_m, i2=c_lo, i3=c_lo + sg.count, i4=slot)
                scalars.setdefault(int(sg.cols[0]), []).append((slot, g))
            else:
                ranges.setdefault((int(sg.cols[0]), sg.count), []).append((("row", row), p.L_m, c_lo, g))
        pools = self.owner.pools
        for (dst, count), contrib in ranges.items():
            flat = []
            rec_rows = {}
            for n, (row, lm, c_lo, g) in enumerate(contrib):
                rec_rows[n] = row[1]
                flat += [0, lm, c_lo, g]
            rec = self.job(ST_GRAD_RANGE, i0=self.base + dst, i1=count, i3=len(contrib))
            rec["contrib"] = (flat, rec_rows)  # table-row bases are patched in at finalisation
        for dst, contrib in scalars.items():
            flat = [v for pair in contrib for v in pair]
            self.job(ST_GRAD_SCALAR, i0=self.base + dst, i2=pools.int(flat), i3=len(contrib))

    # -- Jacobian / Hessian slot runs
    def _post(self, sg: SysSegment):
        if sg.post is None:
            return 0, 0
        if sg.post[0] == "sigma":
            return 1, 0
        return 2, sg.post[1]

    def _slots(self, segs: list[SysSegment]):
        lo, own = self.owner.lo, self.owner
        dst = self.base
        for sg in segs:
            if sg.count == 0:
                continue
            if sg.kind == "phase":
                self._phase_run(sg, dst)
            elif sg.kind == "sys":
                pk, pi_ = self._post(sg)
                self.job(ST_GENERIC, J_SYS, i0=dst, i1=1, i2=-1, i3=self.sys_slot(sg.sys), i4=pk, i5=pi_)
            elif sg.kind == "scaled":
                pk, pi_ = self._post(sg)
                pi, p = sg.phase, lo.phases[sg.phase]
                if sg.nset == "mid":
                    row = self.row_of(pi, "mid", sg.term, True)
                    self.job(
                        ST_GENERIC, J_SCALED, i0=dst, i1=sg.count, i2=-1, i3=self.sys_slot(sg.sys), i4=pk, i5=pi_,
                        i6=("row", row), i7=p.L_m, i8=lo.low[pi].mid_lo,
                    )
                else:
                    slot = self.scalar_of(pi, sg.nset, sg.term, True)
                    self.job(
                        ST_GENERIC, J_SCALED, F_A_SCALAR, i0=dst, i1=1, i2=-1, i3=self.sys_slot(sg.sys),
                        i4=pk, i5=pi_, i6=slot,
                    )
            else:  # outer / tril
                pk, pi_ = self._post(sg)
                a, b = sg.pair
                fa, ia = self._list_source(a)
                fb, ib = self._list_source(b)
                flags = {0: 0, 1: F_A_SCALAR, 2: F_A_UNIT}[fa] | {0: 0, 1: F_B_SCALAR, 2: F_B_UNIT}[fb]
                self.job(
                    ST_GENERIC, J_OUTER if sg.kind == "outer" else J_TRIL, flags, i0=dst, i1=sg.count, i2=-1,
                    i3=self.sys_slot(sg.sys), i4=pk, i5=pi_, i6=ia[0], i7=ia[1], i8=ia[2],
                    i9=ib[0], i10=ib[1], i11=ib[2], i12=b.count,
                )
            dst += sg.count

    def _list_source(self, sl: SysList):
        """(flag bits, (source, L_m, c_lo)); flag 1 = scalar slot, 2 = the constant 1."""
        lo = self.owner.lo
        if sl.term is None:
            return 2, (0, 0, 0)
        pi, p = sl.phase, lo.phases[sl.phase]
        if sl.nset != "mid":
            return 1, (self.scalar_of(pi, sl.nset, sl.term, True), 0, 0)
        row = self.row_of(pi, "mid", sl.term, True)
        c_lo = lo.low[pi].mid_lo
        if sl.summed:
            key = ("sum", pi, row)
            if key not in self.prog[pi].s_store:
                slot = self.new_slot()
                self.prog[pi].s_store[key] = slot
                self.job(ST_REDUCE, i0=("row", row), i1=p.L_m, i2=c_lo, i3=c_lo + lo.low[pi].n_mid, i4=slot)
            return 1, (self.prog[pi].s_store[key], 0, 0)
        return 0, (("row", row), p.L_m, c_lo)

    def _phase_run(self, sg: SysSegment, dst: int):
        lo, own = self.owner.lo, self.owner
        seg: Segment = sg.seg
        pi, p = sg.phase, lo.phases[sg.phase]
        col = p.col
        pools = own.pools
        has_lam = self.sub == HESS
        if seg.kind == "const":
            self.job(ST_GENERIC, J_CONST, i0=dst, i1=seg.count, i2=-1, i3=-1, i6=pools.dbl(seg.data))
        elif seg.kind == "kron":
            nb = len(seg.terms)
            first = self.new_slot(nb)
            for n, t in enumerate(seg.terms):
                self._track(pi, t)
                self.prog[pi].s_store[(seg.nset, term_code(t), first + n)] = first + n
            flags = F_LAM if has_lam else 0
            rows_off = pools.int(seg.lam_rows + sg.lam_base) if has_lam else 0
            self.job(
                ST_GENERIC, J_KRON, flags, f0=seg.sign, i0=dst, i1=seg.count, i2=-1, i3=-1,
                i6=pools.dbl(seg.data), i7=nb, i8=first, i9=rows_off,
            )
        elif seg.kind == "expand" and own.walk_tables[pi] is not None:
            t = seg.terms[0]
            self._track(pi, t)
            lam = sg.lam_base + seg.lam_off if has_lam else -1
            self.prog[pi].walks.append((term_code(t), dst, lam, seg.sign))
        elif seg.kind == "expand":
            row = self.row_of(pi, "mid", seg.terms[0])
            lam = sg.lam_base + seg.lam_off if has_lam else -1
            for piece in own.expand_pieces[pi]:
                if piece["kind"] == "table":
                    self.job(
                        ST_GENERIC, J_EXPAND_TABLE, F_LAM if has_lam else 0, f0=seg.sign,
                        i0=dst + piece["k0"], i1=piece["count"], i2=lam, i3=-1,
                        i6=piece["row"], i7=piece["col"], i8=piece["data"], i9=("row", row), i10=col.L_m,
                    )
                else:
                    # all lists of one state share geometry and multiplier rows: one job, a list table
                    key = (pi, lam, seg.sign, piece["k0"])
                    rec = self.expand_groups.get(key)
                    if rec is None:
                        rec = self.job(
                            ST_EXPAND, 0, F_LAM if has_lam else 0, f0=seg.sign,
                            i2=(lam + piece["row0"]) if has_lam else -1,
                            i3=piece["n"], i4=piece["rows"], i5=piece["step"], i6=piece["c0"],
                            i7=piece["unit"], i8=piece["width"], i10=col.L_m,
                            i11=piece["count"] // piece["rows"],
                        )
                        rec["lists"] = []
                        self.expand_groups[key] = rec
                    rec["lists"].append((dst + piece["k0"], row))
        else:  # direct
            lam = sg.lam_base + seg.lam_off if has_lam else -1
            t = seg.terms[0]
            self._track(pi, t)
            c_lo = {"front": 0, "mid": lo.low[pi].mid_lo, "back": col.L_m - 1}[seg.nset]
            self.prog[pi].direct.append((seg.nset, term_code(t), dst, lam, c_lo))


class DevicePlan:
    """Pools + per-mode plans for one lowered system."""

    def __init__(self, lo: SystemLowering, batch: int = 1, fastmath: bool = False, fused: bool = False,
                 shard: Optional[tuple] = None, node_groups: int = 1, set_subs: tuple = SET_ORDER):
        self.lo = lo
        # callbacks the set pipeline (mode SET) covers, in SET_ORDER
        import os

        self.batch_tables = os.environ.get("POCKIT_B200_BATCH_TABLES", "0") == "1"
        self.set_subs = tuple(m for m in SET_ORDER if m in set(set_subs))
        if not self.set_subs:
            raise ValueError("set_subs must name at least one callback")
        self.B = int(batch)
        self.fastmath = fastmath
        if shard is not None:
            g, G = int(shard[0]), int(shard[1])
            if not 0 <= g < G:
                raise ValueError("shard must be (rank, world) with 0 <= rank < world")
            if fused:
                raise ValueError("mesh sharding is not available with the fused expansion variant")
            weights = np.ones(G) if len(shard) < 3 or shard[2] is None else np.asarray(shard[2], dtype=np.float64)
            if weights.shape != (G,) or not np.all(np.isfinite(weights)) or np.any(weights <= 0):
                raise ValueError("shard weights must be `world` positive numbers")
            # cumulative share boundaries as exact fractions of equal weights when all weights are equal
            if np.all(weights == weights[0]):
                self.shard_cum = [r / G for r in range(G + 1)]
            else:
                c = np.concatenate([[0.0], np.cumsum(weights / weights.sum())])
                c[-1] = 1.0
                self.shard_cum = [float(v) for v in c]
            shard = (g, G)
        self.shard = shard
        # node_groups > 1: up to that many threads per node, one per group of functions.  Measured on B200
        # (round 2): slower -- robot_arm set 59.8 -> 60.0 / 61.4 us with 2 / 4 groups, humanoid 143.7 ->
        # 159.9 / 164.6: re-evaluating the subexpressions the functions share costs more than the shorter
        # per-thread chains save.  Kept as an opt-in switch (POCKIT_B200_NODE_GROUPS).
        self.node_groups = max(1, int(node_groups))
        # fused=True: the per-node program itself walks its block column and writes the slots (no
        # node-table round trip, one launch less).  Measured on B200 (round 1) it is SLOWER than the
        # node program + persistent pk_expand_blocks pair (robot_arm Hessian 40 us vs 31 us, humanoid
        # 100k nodes 254 us vs 159 us per set): one thread per node is too little parallelism for
        # hundreds of serial stores.  Kept selectable (and tested) as the base for a multi-thread-per-node version.
        self.fused = bool(fused)
        self.pools = Pools()
        # scalar-table header: per phase [dt, front values (n_x), back values (n_x)]
        self.header = []
        off = 0
        for p in lo.phases:
            self.header.append(off)
            off += 1 + 2 * p.n_x
        self.header_slots = off
        # FIXED boundary values, per instance: per phase [x0 (n_x), xf (n_x), t0, tf]
        self.fix_off = []
        fix = []
        for p in lo.phases:
            self.fix_off.append(len(fix))
            for info in list(p.info_bc_0) + list(p.info_bc_f) + [p.info_t_0, p.info_t_f]:
                fix.append(float(info.v) if info.t == BcType.FIXED else 0.0)
        self.fixed_default = np.array(fix, dtype=np.float64)
        self.n_fixed = len(fix)
        # mesh tables
        self.tm_off = [self.pools.dbl(p.col.t_m) for p in lo.phases]
        self.wm_off = [self.pools.dbl(p.col.w_m) for p in lo.phases]
        self.defect_tables = [self._defect_tables(p) for p in lo.phases]
        self.expand_pieces = [self._expand_pieces(p) for p in lo.phases]
        self.walk_tables = [self._walk_tables(p) if self.fused else None for p in lo.phases]
        self.modes: dict[int, ModePlan] = {}

    def mode(self, m: int) -> ModePlan:
        if m not in self.modes:
            self.modes[m] = ModePlan(self, m)
        return self.modes[m]

    # ---------------------------------------------------------------- constant tables
    def _defect_tables(self, p):
        col = p.col
        T, I = col.T, col.I
        rows = np.concatenate([I.f.row, I.m.row, I.b.row]).astype(np.int64)
        cols = np.concatenate([I.f.col, I.m.col, I.b.col]).astype(np.int64)
        data = np.concatenate([I.f.data, I.m.data, I.b.data])
        order = np.concatenate([I.f.k, I.m.k, I.b.k]).argsort(kind="stable")
        rows, cols, data = rows[order], cols[order], data[order]
        row_ptr = np.searchsorted(rows, np.arange(col.n_rows + 1))
        trow = np.concatenate([T.f.row, T.m.row, T.b.row]).astype(np.int64)
        tcol = np.concatenate([T.f.col, T.m.col, T.b.col]).astype(np.int64)
        tval = np.concatenate([T.f.data, T.m.data, T.b.data])
        tpos = np.zeros(col.n_rows, dtype=np.int64)
        tneg = np.zeros(col.n_rows, dtype=np.int64)
        tpos[trow[tval > 0]] = tcol[tval > 0]
        tneg[trow[tval < 0]] = tcol[tval < 0]
        P = self.pools
        fast = dict(n=0, rows_blk=0, unit=0, width=0)
        if col.same_order and col.dense_blocks:
            from .discretization import _unit_integration_block

            n = int(col.num_point[0])
            unit = _unit_integration_block(col.scheme, n)
            # column-major for the defect kernel: lanes own consecutive rows and read one column at a time
            fast = dict(n=n, rows_blk=unit.shape[0], unit=P.dbl(np.ascontiguousarray(unit.T).ravel()), width=P.dbl(col.width))
        return dict(
            row_ptr=P.int(row_ptr), col=P.int(cols), data=P.dbl(data), tpos=P.int(tpos), tneg=P.int(tneg), **fast
        )

    def _walk_tables(self, p):
        """Per-interval geometry for the FUSED expansion (the per-node program itself walks down
        its column of the interval block and writes the slots): for interval K
        ``[n, rows, row length of its middle part, first kept column, first node, first triplet,
        dpool offset of the unit block, first defect row]`` plus the node -> interval map.
        ``None`` when a block lost an exact zero (slot arithmetic would be off): table path."""
        from .discretization import _unit_integration_block

        col = p.col
        if not col.dense_blocks or p.n_x == 0:
            return None
        Im = col.I.m
        nK = len(col.num_point)
        lgl = col.scheme == "lgl"
        off = np.searchsorted(Im.row, col.row_start)
        rec = np.zeros((nK, 8), dtype=np.int64)
        node_iv = np.zeros(col.L_m, dtype=np.int64)
        for K in range(nK):
            n = int(col.num_point[K])
            rows = n - 1 if lgl else n
            first = 1 if K == 0 else 0
            length = n - first - (1 if lgl and K == nK - 1 else 0)
            if rows * length != off[K + 1] - off[K]:
                return None
            unit = _unit_integration_block(col.scheme, n)
            rec[K] = [n, rows, length, first, col.l_m[K], off[K], self.pools.dbl(unit.ravel()), col.row_start[K]]
            node_iv[col.l_m[K] : col.r_m[K]] = K  # LGL border nodes end up in the interval they start
        return dict(node_iv=self.pools.int(node_iv), rec=self.pools.int(rec.ravel()), width=self.pools.dbl(col.width),
                    lgl=lgl)

    def _expand_pieces(self, p):
        """Split the middle part of the integration operator into pieces the EXPAND kernel can
        address arithmetically -- maximal runs of complete interval blocks of equal order (an
        hp-refined mesh gives several runs) -- and table-driven remainders (the first interval
        loses its front column, the last LGL interval its back column, and blocks that lost an
        exact zero are irregular)."""
        from .discretization import _unit_integration_block

        col = p.col
        Im = col.I.m
        P = self.pools
        nK = len(col.num_point)
        lgl = col.scheme == "lgl"
        npt = col.num_point.astype(np.int64)
        off = np.searchsorted(Im.row, col.row_start)  # first triplet of every interval (+ end)
        table = lambda k0, k1: dict(
            kind="table", k0=int(k0), count=int(k1 - k0), row=P.int(Im.row[k0:k1]), col=P.int(Im.col[k0:k1]),
            data=P.dbl(Im.data[k0:k1]),
        )
        width_off = P.dbl(col.width)
        pieces, pending = [], None  # pending = start of the current table stretch

        def flush(k_end):
            nonlocal pending
            if pending is not None and k_end > pending:
                pieces.append(table(pending, k_end))
            pending = None

        # Batches of small problems (BASELINE configs[4]: 8192 instances of a 14-interval mesh): the block
        # kernels would give every instance its own short, latency-bound block and walk 5 x 6 blocks in
        # 48-byte runs; the table-driven expansion of pk_generic_jobs instead flattens (instance, slot),
        # writes fully coalesced and reads its three small per-slot tables (shared by all instances)
        # through L1.  Used when an instance has fewer (interval, column) pairs than the parameter-driven
        # kernel needs (1024).  Opt-in (POCKIT_B200_BATCH_TABLES=1): measured on B200 in round 2
        # (profiles/r02_call10_small_kernels.log) the Jacobian stage is shorter (91 vs 109 us) but the whole
        # batched set is slower (301 vs 281 us) -- three 8-byte table reads per slot load the L1 path that the
        # other callbacks' kernels also need.
        table_only = self.B > 1 and int(npt.sum()) < 1024 and self.batch_tables
        K = 0
        while K < nK:
            n = int(npt[K])
            rows = n - 1 if lgl else n
            complete = (K > 0 and not (lgl and K == nK - 1) and off[K + 1] - off[K] == rows * n and n <= 128
                        and not table_only)
            if not complete:
                if pending is None:
                    pending = int(off[K])
                K += 1
                continue
            Kb = K
            while (
                Kb < nK and npt[Kb] == n and not (lgl and Kb == nK - 1) and off[Kb + 1] - off[Kb] == rows * n
            ):
                Kb += 1
            flush(int(off[K]))
            unit = _unit_integration_block(col.scheme, n)
            pieces.append(
                dict(
                    kind="block", k0=int(off[K]), count=int(off[Kb] - off[K]), n=n, rows=rows, step=rows,
                    c0=int(col.l_m[K]), row0=int(col.row_start[K]), unit=P.dbl(unit.ravel()), width=width_off + K,
                )
            )
            K = Kb
        flush(len(Im))
        return [q for q in pieces if q["count"]]

    # ---------------------------------------------------------------- code generation
    def source(self, m: int) -> tuple[str, list[str], Optional[str]]:
        """CUDA C of mode ``m``: (source, node kernel names, system kernel name or None)."""
        mp = self.mode(m)
        if mp.src is not None:
            return mp.src
        lo = self.lo
        name = MODES[m]
        body, kernels = [], []
        for pi in range(len(lo.phases)):
            kname = f"pk_node_{name}_p{pi}"
            body.append(self._node_kernel(mp, pi, kname))
            kernels.append(kname)
        sys_name = None
        if mp.sys_need or any(w.any() for w in mp.int_needed):
            sys_name = f"pk_sys_{name}"
            body.append(self._sys_kernel(mp, sys_name))
        head = [
            "// generated by pockit_b200.plan -- per-node programs of mode '%s'" % name,
            PRELUDE,
            f"__constant__ long long pk_tab_{name}[{max(1, len(mp.table))}];",
        ]
        mp.src = ("\n".join(head + body) + "\n", kernels, sys_name)
        return mp.src

    def finalize(self, m: int) -> dict:
        """Everything the engine needs for mode ``m`` (call after all modes were
        built so the pools are complete): source, table, resolved job arrays."""
        mp = self.mode(m)
        src, kernels, sys_name = self.source(m)
        jobs = {}
        for stage, recs in mp.jobs.items():
            arr = np.zeros(len(recs), dtype=JOB_DTYPE)
            for n, rec in enumerate(recs):
                iv = list(rec["i"])
                for idx, row in rec["rows"].items():
                    iv[idx] = mp.row_base(row)
                if "contrib" in rec:
                    flat, rows = rec["contrib"]
                    flat = list(flat)
                    for cn, row in rows.items():
                        flat[4 * cn] = mp.row_base(row)
                    iv[2] = self.pools.int(flat)
                if "lists" in rec:
                    flat = [v for dst_, row in rec["lists"] for v in (dst_, mp.row_base(row))]
                    iv[0] = self.pools.int(flat)
                    iv[1] = len(rec["lists"])
                arr[n] = (rec["type"], rec["flags"], iv, rec["f"])
            jobs[stage] = arr
        return dict(
            source=src, kernels=kernels, sys_kernel=sys_name, table=np.array(mp.table, dtype=np.int64),
            table_symbol=f"pk_tab_{MODES[m]}", jobs=jobs, n_scalar=mp.n_scalar, n_out=mp.n_out,
            n_table=mp.n_table, runs=None if mp.owned_runs is None else np.array(mp.owned_runs, dtype=np.int64).reshape(-1, 2),
            grad_range=mp.grad_range, sub_range=dict(mp.sub_range), node_threads=list(mp.node_threads),
        )

    def _function_groups(self, mp: ModePlan, pi: int) -> list:
        """Partition the phase's functions ``(family, index)`` that this mode stores something of
        into at most ``node_groups`` groups of similar cost (operation count of the needed leaves),
        largest first onto the lightest group.  One group (everything) unless asked otherwise."""
        prog = mp.prog[pi]
        p = self.lo.phases[pi]
        fam = {"d": p.F_d, "I": p.F_I, "c": p.F_c}
        cost: dict = {}
        for lf in prog.need:
            fn = fam[lf.key[0]][lf.key[1]]
            e = fn.expr if lf.kind == "F" else (fn.G_expr if lf.kind == "G" else fn.H_expr)[lf.key[2]]
            cost[(lf.key[0], lf.key[1])] = cost.get((lf.key[0], lf.key[1]), 0) + 1 + int(sp.count_ops(e))
        k = max(1, min(int(self.node_groups), len(cost))) if not self.fused else 1
        if k <= 1:
            return [set(cost)]
        groups, load = [set() for _ in range(k)], [0] * k
        for fn_key, c in sorted(cost.items(), key=lambda kv: (-kv[1], kv[0])):
            g = load.index(min(load))
            groups[g].add(fn_key)
            load[g] += c
        return [g for g in groups if g] or [set()]

    def _node_kernel(self, mp: ModePlan, pi: int, kname: str) -> str:
        lo = self.lo
        p = lo.phases[pi]
        col = p.col
        prog = mp.prog[pi]
        name = MODES[mp.mode]
        T = lambda v: f"T[{mp.tab(v)}]"
        has_back = col.index_mstage.b
        n_x, n_u, n_s = p.n_x, p.n_u, p.n_s
        L = []
        A = L.append
        A(f'extern "C" __global__ void __launch_bounds__({NODE_BLOCK}) {kname}(')
        A("    const double* __restrict__ X, const double* __restrict__ LAM, const double* __restrict__ FIX,")
        A("    const double* __restrict__ TM, const double* __restrict__ WM,")
        A("    double* __restrict__ S, double* __restrict__ W, double* __restrict__ OUT, int B,")
        A("    const double* __restrict__ DP, const long long* __restrict__ IP)")
        A("{")
        A(f"    const long long* T = pk_tab_{name};")
        A(f"    const int Lm = (int){T(col.L_m)};")
        # expression groups (opt-in, DevicePlan(node_groups=k)): the functions of the phase are spread
        # over up to k threads per node, each with the leaves (and the CSE) of its own functions only --
        # "one thread per node per expression group".  Every store term carries leaves of exactly one
        # function, so the stores split cleanly; group-major thread order keeps warps uniform.
        fn_groups = self._function_groups(mp, pi)
        G = len(fn_groups)
        mp.node_threads[pi] = G
        if G == 1:
            A("    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;")
            A("    if (gid >= (long long)B * Lm) return;")
            A("    const int grp = 0;")
        else:
            A("    const long long gid0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;")
            A(f"    if (gid0 >= (long long)B * Lm * {G}) return;")
            A("    const int grp = (int)(gid0 / ((long long)B * Lm));")
            A("    const long long gid = gid0 - (long long)grp * B * Lm;")
        A("    const int b = (int)(gid / Lm);")
        A("    const int c = (int)(gid - (long long)b * Lm);")
        A(f"    const double* xs = X + (long long)b * {T(lo.r_s)};")
        A(f"    const double* xp = xs + {T(lo.l_p[pi])};")
        A(f"    const double* sv = xs + {T(lo.l_s)};")
        A(f"    double* Sb = S + (long long)b * {T(mp.n_scalar_final)};")
        A(f"    const double* lam = LAM + (long long)b * {T(lo.m)};")
        A(f"    double* out = OUT + (long long)b * {T(mp.n_out)};")
        A(f"    const double* fix = FIX + (long long)b * {T(self.n_fixed)} + {T(self.fix_off[pi])};")
        A(f"    const int Lx = (int){T(col.L_x)};")
        A("    const bool first = (c == 0), last = (c == Lm - 1);")
        names = {}
        for k, sym in enumerate(p.s):
            A(f"    const double s{k} = sv[{k}];")
            names[sym] = f"s{k}"
        # boundary functions of s: values and derivative leaves
        bnd = []
        infos = (
            [(("x0", i), p.info_bc_0[i]) for i in range(n_x)]
            + [(("xf", i), p.info_bc_f[i]) for i in range(n_x)]
            + [(("t0",), p.info_t_0), (("tf",), p.info_t_f)]
        )
        for tag, info in infos:
            if info.t == BcType.FUNC:
                tg = "_".join(str(v) for v in tag)
                bnd.append((f"bV_{tg}", info.v.expr))
                bnd += [(f"bG_{tg}_{jj}", e) for jj, e in enumerate(info.v.G_expr)]
                bnd += [(f"bH_{tg}_{jj}", e) for jj, e in enumerate(info.v.H_expr)]
        if bnd:
            L.append(emit_block(bnd, names, prefix="cb").rstrip("\n"))

        def bc_value(tag, info, free_expr, fix_index):
            if info.t == BcType.FREE:
                return free_expr
            if info.t == BcType.FIXED:
                return f"fix[{fix_index}]"
            return "bV_" + "_".join(str(v) for v in tag)

        A(f"    const double t0 = {bc_value(('t0',), p.info_t_0, f'xp[{T(col.L - 2)}]', 2 * n_x)};")
        A(f"    const double tf = {bc_value(('tf',), p.info_t_f, f'xp[{T(col.L - 1)}]', 2 * n_x + 1)};")
        A("    const double dt = tf - t0;")
        A("    const double mt = (tf + t0) / 2.0;")
        A("    const double tm = TM[c];")
        A("    const double om = 1.0 - tm;")
        A("    const double wq = WM[c];")
        A("    const double t = (tm - 0.5) * dt + mt;")
        names[p.t] = "t"
        for i, sym in enumerate(p.x):
            A(f"    double x{i} = xp[{i}LL * Lx + c];")
            if p.info_bc_0[i].t != BcType.FREE:
                A(f"    if (first) x{i} = {bc_value(('x0', i), p.info_bc_0[i], '', i)};")
            if has_back and p.info_bc_f[i].t != BcType.FREE:
                A(f"    if (last) x{i} = {bc_value(('xf', i), p.info_bc_f[i], '', n_x + i)};")
            names[sym] = f"x{i}"
        for j, sym in enumerate(p.u):
            A(f"    const double u{j} = xp[{n_x}LL * Lx + {j}LL * Lm + c];")
            names[sym] = f"u{j}"
        # function leaves needed by this mode: one CSE per expression group
        fam = {"d": p.F_d, "I": p.F_I, "c": p.F_c}
        A("    const long long nd = (long long)b * Lm + c;")

        def group_of(code):
            fn = _function_of(code)
            for g, members in enumerate(fn_groups):
                if fn in members:
                    return g
            return 0

        def stores(nset, indent, grp):
            body = []
            for key, slot in prog.s_store.items():
                if key[0] == nset and group_of(key[1]) == grp:
                    body.append(f"{indent}Sb[{T(slot)}] = {key[1]};")
            for key, row in prog.w_store.items():
                if key[0] == nset and group_of(key[1]) == grp:
                    body.append(f"{indent}W[{T(('row', row))} + nd] = {key[1]};")
            for ns, code, dst, lam, c_lo in prog.direct:
                if ns == nset and group_of(code) == grp:
                    e = f"c - {T(c_lo)}"
                    val = code if lam < 0 else f"{code} * lam[{T(lam)} + {e}]"
                    body.append(f"{indent}out[{T(dst)} + {e}] = {val};")
            return body

        for grp, members in enumerate(fn_groups):
            A(f"    if (grp == {grp}) {{")
            leaves = []
            for lf in sorted(prog.need, key=lambda l: (l.key[0], l.key[1], l.kind, l.key[2:] or (0,))):
                if (lf.key[0], lf.key[1]) not in members:
                    continue
                fn = fam[lf.key[0]][lf.key[1]]
                e = fn.expr if lf.kind == "F" else (fn.G_expr if lf.kind == "G" else fn.H_expr)[lf.key[2]]
                leaves.append((_ident(lf), e))
            if leaves:
                L.append(emit_block(leaves, names, prefix=f"ce{grp}" if G > 1 else "ce").rstrip("\n"))
            L += stores("all", "    ", grp)
            A("    if (first) {")
            if grp == 0:
                A(f"        Sb[{T(self.header[pi])}] = dt;")
                for i in range(n_x):
                    A(f"        Sb[{T(self.header[pi] + 1 + i)}] = x{i};")
            L += stores("basic", "        ", grp)
            L += stores("front", "        ", grp)
            if has_back:
                A("    } else if (last) {")
                L += stores("back", "        ", grp)
            A("    } else {")
            L += stores("mid", "        ", grp)
            if prog.walks and grp == 0:
                wt = self.walk_tables[pi]
                A("        // fused expansion through the integration operator: this node's column of its")
                A("        // interval block (two columns for a shared LGL border node), phasebase.py:1120-1124, 1280-1285")
                A(f"        const int K = (int)IP[{T(wt['node_iv'])} + c];")
                A(f"        const long long* rk = IP + {T(wt['rec'])} + 8LL * K;")
                A("        const int cc = c - (int)rk[4];")
                A(f"        const PkColumn col0 = pk_column(DP, rk, cc, DP[{T(wt['width'])} + K]);")
                if wt["lgl"]:
                    A("        const bool shared = (cc == 0 && K > 0);")
                    A("        const long long* rj = shared ? rk - 8 : rk;")
                    A(f"        const PkColumn col1 = pk_column(DP, rj, (int)rj[0] - 1, DP[{T(wt['width'])} + (shared ? K - 1 : K)]);")
                for code, dst, lam, sign in prog.walks:
                    lam0 = "nullptr" if lam < 0 else f"lam + {T(lam)}"
                    A(f"        {{ const double v = {code}; double* o = out + {T(dst)}; const double* lm = {lam0};")
                    A(f"          pk_walk(o, v, col0, lm, {sign!r});")
                    if wt["lgl"]:
                        A(f"          if (shared) pk_walk(o, v, col1, lm, {sign!r}); }}")
                    else:
                        A("        }")
            A("    }")
            A("    }")
        A("    if (grp != 0) return;")
        A("    if (last) {")
        for i in range(n_x):
            if has_back:
                A(f"        Sb[{T(self.header[pi] + 1 + n_x + i)}] = x{i};")
            else:
                free = f"xp[{i}LL * Lx + Lm]"
                A(
                    f"        Sb[{T(self.header[pi] + 1 + n_x + i)}] = "
                    f"{bc_value(('xf', i), p.info_bc_f[i], free, n_x + i)};"
                )
        A("    }")
        A("}")
        return "\n".join(L)

    # ---------------------------------------------------------------- continuous error estimate
    def error_estimate(self) -> dict:
        """Programs and operators of the continuous error-estimate data (SURVEY 8f.3,
        ``phasebase.py:1339-1366``): per phase a *prep* program (the phase vector with its FIXED /
        FUNC boundary values substituted), a *node* program (the dynamics at the augmented-mesh
        nodes, read from the interpolated states / controls) and the CSR operators ``V`` (mesh ->
        augmented mesh), ``T`` (translation) and ``I`` (augmented integration) the engine applies
        with sequential row sums, like the reference's ``csr.dot``."""
        from .discretization import AugmentedCollocation

        lo = self.lo
        body, phases = [], []
        for pi, p in enumerate(lo.phases):
            col = p.col
            A = AugmentedCollocation(col)
            n_x, n_u = p.n_x, p.n_u
            names = {}
            L = []
            add = L.append
            prep, node = f"pk_aug_prep_p{pi}", f"pk_aug_node_p{pi}"
            # --- prep: xs = phase slice of x with boundary values substituted (phasebase.py:1340-1347)
            add(f'extern "C" __global__ void {prep}(const double* __restrict__ X, const double* __restrict__ FIX,')
            add("    double* __restrict__ XS, int B)")
            add("{")
            add("    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;")
            add(f"    if (gid >= (long long)B * {col.L}) return;")
            add(f"    const int b = (int)(gid / {col.L});")
            add(f"    const int k = (int)(gid - (long long)b * {col.L});")
            add(f"    const double* xp = X + (long long)b * {lo.r_s} + {int(lo.l_p[pi])};")
            add(f"    const double* sv = X + (long long)b * {lo.r_s} + {lo.l_s};")
            add(f"    const double* fix = FIX + (long long)b * {self.n_fixed} + {self.fix_off[pi]};")
            for k, sym in enumerate(p.s):
                add(f"    const double s{k} = sv[{k}];")
                names[sym] = f"s{k}"
            infos = ([(("x0", i), p.info_bc_0[i], int(col.l_v[i]), i) for i in range(n_x)]
                     + [(("xf", i), p.info_bc_f[i], int(col.r_v[i]) - 1, n_x + i) for i in range(n_x)]
                     + [(("t0",), p.info_t_0, col.L - 2, 2 * n_x), (("tf",), p.info_t_f, col.L - 1, 2 * n_x + 1)])
            bnd = [("bV_" + "_".join(str(v) for v in tag), info.v.expr) for tag, info, _, _ in infos if info.t == BcType.FUNC]
            if bnd:
                L.append(emit_block(bnd, names, prefix="cb").rstrip("\n"))
            add("    double v = xp[k];")
            for tag, info, slot, fix_index in infos:
                if info.t == BcType.FIXED:
                    add(f"    if (k == {slot}) v = fix[{fix_index}];")
                elif info.t == BcType.FUNC:
                    add(f"    if (k == {slot}) v = bV_{'_'.join(str(q) for q in tag)};")
            add(f"    XS[(long long)b * {col.L} + k] = v;")
            add("}")
            # --- node: dynamics at the augmented nodes
            add(f'extern "C" __global__ void {node}(const double* __restrict__ X, const double* __restrict__ XS,')
            add("    const double* __restrict__ XU, const double* __restrict__ TMA, double* __restrict__ WA, int B)")
            add("{")
            add("    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;")
            add(f"    if (gid >= (long long)B * {A.L_m}) return;")
            add(f"    const int b = (int)(gid / {A.L_m});")
            add(f"    const int c = (int)(gid - (long long)b * {A.L_m});")
            add(f"    const double* sv = X + (long long)b * {lo.r_s} + {lo.l_s};")
            names = {}
            for k, sym in enumerate(p.s):
                add(f"    const double s{k} = sv[{k}];")
                names[sym] = f"s{k}"
            add(f"    const double* xsb = XS + (long long)b * {col.L};")
            add(f"    const double t0 = xsb[{col.L - 2}], tf = xsb[{col.L - 1}];")
            add("    const double dt = tf - t0;")
            add("    const double mt = (tf + t0) / 2.0;")
            add("    const double t = (TMA[c] - 0.5) * dt + mt;")
            names[p.t] = "t"
            add(f"    const double* xu = XU + (long long)b * {(n_x + n_u) * A.L_m};")
            for i, sym in enumerate(p.x):
                add(f"    const double x{i} = xu[{i * A.L_m}LL + c];")
                names[sym] = f"x{i}"
            for j, sym in enumerate(p.u):
                add(f"    const double u{j} = xu[{(n_x + j) * A.L_m}LL + c];")
                names[sym] = f"u{j}"
            leaves = [(f"F_d_{i}", p.F_d[i].expr) for i in range(n_x)]
            if leaves:
                L.append(emit_block(leaves, names, prefix="ce").rstrip("\n"))
            for i in range(n_x):
                add(f"    WA[((long long)b * {n_x} + {i}) * {A.L_m} + c] = F_d_{i};")
            add("}")
            body.append("\n".join(L))
            phases.append(dict(prep=prep, node=node, x_offset=int(lo.l_p[pi]), L=int(col.L), L_xu=int(col.L_xu),
                               L_x_all=int(col.r_v[n_x - 1]) if n_x else 0, n_x=n_x, n_u=n_u, Lm_aug=int(A.L_m),
                               rows=int(A.rows), tm_aug=np.ascontiguousarray(A.t_m), V=A.V, T=A.T, I=A.I, aug=A))
        head = ["// generated by pockit_b200.plan -- continuous error-estimate programs", PRELUDE]
        return dict(source="\n".join(head + body) + "\n", phases=phases)

    def _sys_kernel(self, mp: ModePlan, kname: str) -> str:
        lo = self.lo
        name = MODES[mp.mode]
        T = lambda v: f"T[{mp.tab(v)}]"
        L = []
        A = L.append
        A(f'extern "C" __global__ void {kname}(')
        A("    const double* __restrict__ X, double* __restrict__ S, double* __restrict__ OUT, int B)")
        A("{")
        A(f"    const long long* T = pk_tab_{name};")
        A("    const int b = blockIdx.x * blockDim.x + threadIdx.x;")
        A("    if (b >= B) return;")
        A(f"    const double* sv = X + (long long)b * {T(lo.r_s)} + {T(lo.l_s)};")
        A(f"    double* Sb = S + (long long)b * {T(mp.n_scalar_final)};")
        A(f"    double* out = OUT + (long long)b * {T(mp.n_out)};")
        names = {}
        n = 0
        for pi, p in enumerate(lo.phases):
            for k, sym in enumerate(p.I):
                if mp.int_needed[pi][k]:
                    # vi.dot(w_m) * dt   (phasebase.py:1004)
                    A(f"    const double I{n} = Sb[{T(mp.int_sum_slot[(pi, k)])}] * Sb[{T(self.header[pi])}];")
                else:
                    A(f"    const double I{n} = 0.0;")
                names[sym] = f"I{n}"
                n += 1
        for k, sym in enumerate(lo.system.s):
            A(f"    const double s{k} = sv[{k}];")
            names[sym] = f"s{k}"
        outs = []
        for lf in mp.sys_need:
            fn = lo.F_o if lf.key[0] == "o" else lo.F_c[lf.key[1]]
            idx = lf.key[-1]
            e = fn.expr if lf.kind == "sF" else (fn.G_expr if lf.kind == "sG" else fn.H_expr)[idx]
            outs.append((_ident(lf), e))
        if outs:
            L.append(emit_block(outs, names, prefix="cs").rstrip("\n"))
        for lf, slot in mp.sys_need.items():
            if slot < 0:
                A(f"    out[{-1 - slot}] = {_ident(lf)};")
            else:
                A(f"    Sb[{T(slot)}] = {_ident(lf)};")
        A("}")
        return "\n".join(L)
