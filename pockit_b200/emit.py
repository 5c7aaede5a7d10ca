"""SymPy -> CUDA C.

The reference turns each expression into NumPy source text and lets Numba compile
it (``pockit/base/fastfunc.py:41-43, 271-308``).  Here the expression *trees*
are printed as double-precision CUDA C statements (own printer, no regex on
Python text), with one common-subexpression pass shared by a function's value
and all of its derivatives so that transcendental calls are evaluated once
per node.  Small integer powers are expanded to products like the reference's
``create_expand_pow_optimization(3)`` (:180).
"""
from __future__ import annotations

import sympy as sp

__all__ = ["CExpr", "emit_block", "PRELUDE"]

PRELUDE = r"""
__device__ __forceinline__ double pk_sq(double v) { return v * v; }
__device__ __forceinline__ double pk_cube(double v) { return v * v * v; }
__device__ __forceinline__ double pk_sign(double v) { return (v > 0.0) - (v < 0.0); }
__device__ __forceinline__ double pk_heaviside(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? 0.0 : 0.5); }

// One column of one interval block of the integration operator, as seen by the node that owns it:
// entries (unit[r][c] * width) / 2 for r = 0..rows-1 land `stride` slots apart, starting at `off`.
struct PkColumn { const double* u; double w; long long off; long long row0; int n, rows, stride; };
__device__ __forceinline__ PkColumn pk_column(const double* DP, const long long* rec, int cc, double width) {
    PkColumn k;
    k.n = (int)rec[0]; k.rows = (int)rec[1]; k.stride = (int)rec[2];
    k.u = DP + rec[6] + cc; k.w = width;
    k.off = rec[5] + (cc - rec[3]); k.row0 = rec[7];
    return k;
}
// slots of one list in that column: sign * ((unit * w) / 2) [* lam_row] * v, the reference's association
__device__ __forceinline__ void pk_walk(double* __restrict__ o, double v, const PkColumn& k,
                                        const double* __restrict__ lam, double sign) {
    o += k.off;
    if (lam) {
        lam += k.row0;
#pragma unroll 4
        for (int r = 0; r < k.rows; ++r) o[(long long)r * k.stride] = ((((sign * k.u[r * k.n]) * k.w) / 2.0) * lam[r]) * v;
    } else {
#pragma unroll 4
        for (int r = 0; r < k.rows; ++r) o[(long long)r * k.stride] = (((sign * k.u[r * k.n]) * k.w) / 2.0) * v;
    }
}
"""

_UNARY = {
    sp.sin: "sin", sp.cos: "cos", sp.tan: "tan", sp.asin: "asin", sp.acos: "acos", sp.atan: "atan",
    sp.sinh: "sinh", sp.cosh: "cosh", sp.tanh: "tanh", sp.asinh: "asinh", sp.acosh: "acosh",
    sp.atanh: "atanh", sp.exp: "exp", sp.log: "log", sp.Abs: "fabs", sp.sign: "pk_sign",
    sp.floor: "floor", sp.ceiling: "ceil", sp.erf: "erf", sp.erfc: "erfc",
}


def _lit(v: float) -> str:
    v = float(v)
    if v != v:
        return "(0.0/0.0)"
    if v in (float("inf"), float("-inf")):
        return "(1.0/0.0)" if v > 0 else "(-1.0/0.0)"
    r = repr(abs(v))
    if "e" not in r and "." not in r:
        r += ".0"
    return f"(-{r})" if v < 0 or (v == 0 and str(v).startswith("-")) else r


class CExpr:
    """Printer with a symbol -> C identifier map."""

    def __init__(self, names: dict):
        self.names = dict(names)

    def __call__(self, e) -> str:
        return self._p(sp.sympify(e))

    def _p(self, e) -> str:
        if e.is_Symbol:
            return self.names[e]
        if e.is_Integer:
            return _lit(int(e))
        if e.is_Rational:
            return f"({_lit(int(e.p))}/{_lit(int(e.q))})"
        if e.is_Float:
            return _lit(e)
        if e.is_NumberSymbol or e.is_number and e.is_real and not e.args:
            return _lit(e.evalf(17))
        if e.is_Add:
            out = ""
            for i, t in enumerate(e.as_ordered_terms()):
                neg = t.could_extract_minus_sign()
                body = self._p(-t if neg else t)
                if i == 0:
                    out = f"-{body}" if neg else body
                else:
                    out += f" - {body}" if neg else f" + {body}"
            return f"({out})"
        if e.is_Mul:
            num, den = [], []
            negative = False
            for f in e.as_ordered_factors():
                if f.is_Pow and f.exp.is_number and f.exp.could_extract_minus_sign():
                    den.append(sp.Pow(f.base, -f.exp))
                elif f.is_Rational and not f.is_Integer:
                    negative ^= f.p < 0
                    if abs(f.p) != 1:
                        num.append(sp.Integer(abs(f.p)))
                    den.append(sp.Integer(f.q))
                elif f.is_number and f.is_real and f < 0:
                    negative = not negative
                    if f != -1:
                        num.append(-f)
                else:
                    num.append(f)
            n = "*".join(self._p(f) for f in num) if num else "1.0"
            if den:
                n = f"{n}/(" + "*".join(self._p(f) for f in den) + ")"
            return f"(-{n})" if negative else f"({n})"
        if e.is_Pow:
            return self._pow(e.base, e.exp)
        if isinstance(e, sp.Function):
            fn = _UNARY.get(type(e))
            if fn is not None and len(e.args) == 1:
                return f"{fn}({self._p(e.args[0])})"
            if isinstance(e, sp.atan2):
                return f"atan2({self._p(e.args[0])}, {self._p(e.args[1])})"
            if isinstance(e, (sp.Max, sp.Min)):
                op = "fmax" if isinstance(e, sp.Max) else "fmin"
                acc = self._p(e.args[0])
                for a in e.args[1:]:
                    acc = f"{op}({acc}, {self._p(a)})"
                return acc
            if isinstance(e, sp.Heaviside):
                return f"pk_heaviside({self._p(e.args[0])})"
        if isinstance(e, sp.Piecewise):
            out = "(0.0/0.0)"
            for val, cond in reversed(e.args):
                out = self._p(val) if cond == True else f"({self._rel(cond)} ? {self._p(val)} : {out})"  # noqa: E712
            return out
        raise NotImplementedError(f"cannot lower {type(e).__name__}: {e}")

    def _rel(self, c) -> str:
        ops = {sp.StrictLessThan: "<", sp.LessThan: "<=", sp.StrictGreaterThan: ">", sp.GreaterThan: ">=",
               sp.Equality: "==", sp.Unequality: "!="}
        if type(c) in ops:
            return f"({self._p(c.lhs)} {ops[type(c)]} {self._p(c.rhs)})"
        if isinstance(c, sp.And):
            return "(" + " && ".join(self._rel(a) for a in c.args) + ")"
        if isinstance(c, sp.Or):
            return "(" + " || ".join(self._rel(a) for a in c.args) + ")"
        raise NotImplementedError(f"cannot lower condition {c}")

    @staticmethod
    def _par(s: str) -> str:
        return s

    def _pow(self, base, exp) -> str:
        b = self._p(base)
        if exp.is_Integer:
            n = int(exp)
            if n == 1:
                return b
            if n == 2:
                return f"pk_sq({b})"
            if n == 3 and base.is_Symbol:
                return f"pk_cube({b})"
            if n == -1:
                return f"(1.0/{self._par(b)})"
            if n == -2:
                return f"(1.0/pk_sq({b}))"
            if n == -3 and base.is_Symbol:
                return f"(1.0/pk_cube({b}))"
            return f"pow({b}, {_lit(n)})"
        if exp == sp.Rational(1, 2):
            return f"sqrt({b})"
        if exp == sp.Rational(-1, 2):
            return f"(1.0/sqrt({b}))"
        return f"pow({b}, {self._p(exp)})"


def emit_block(outputs: list[tuple[str, sp.Expr]], names: dict, indent: str = "    ", prefix: str = "c") -> str:
    """C statements computing every ``(identifier, expression)`` pair with one
    shared CSE pass.  ``names`` maps free symbols to C identifiers."""
    if not outputs:
        return ""
    exprs = [sp.sympify(e) for _, e in outputs]
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols(f"{prefix}_"), optimizations="basic", order="none")
    pr = CExpr(names)
    lines = []
    for sym, sub in repl:
        pr.names[sym] = str(sym)
        lines.append(f"{indent}const double {sym} = {pr(sub)};")
    for (name, _), e in zip(outputs, red):
        lines.append(f"{indent}const double {name} = {pr(e)};")
    return "\n".join(lines) + "\n"
