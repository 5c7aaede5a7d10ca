"""Sparse symbolic first/second derivatives of one scalar expression.

Device-side counterpart of the reference's ``FastFunc``
(``pockit/base/fastfunc.py:78-317``): it keeps the *same* sparsity contract --
``G_index`` lists the arguments with a non-zero first derivative in argument
order; ``(H_index_row, H_index_col)`` the non-zero lower-triangle second
derivatives, rows outer / columns inner (``fastfunc.py:237-269``) -- but instead
of compiling Numba kernels it only stores the SymPy trees.  They are lowered to
CUDA C by :mod:`pockit_b200.emit` together with everything that consumes them.
"""
from __future__ import annotations

import numpy as np
import sympy as sp

__all__ = ["SymFunc"]


class SymFunc:
    def __init__(self, function, args: list[sp.Symbol], simplify: bool = False):
        expr = sp.sympify(function)
        self.args = list(args)
        if simplify:
            expr = sp.simplify(expr)
        self.expr = expr
        position = {a: i for i, a in enumerate(self.args)}

        def used(e) -> list[int]:
            return sorted(position[s] for s in e.free_symbols)

        self.G_expr: list[sp.Expr] = []
        self.H_expr: list[sp.Expr] = []
        g_index: list[int] = []
        h_row: list[int] = []
        h_col: list[int] = []
        for j in used(expr):
            d1 = sp.diff(expr, self.args[j])
            if simplify:
                d1 = sp.simplify(d1)
            if d1 == 0:
                continue
            self.G_expr.append(d1)
            g_index.append(j)
            for k in used(d1):
                if k > j:
                    break
                d2 = sp.diff(d1, self.args[k])
                if simplify:
                    d2 = sp.simplify(d2)
                if d2 == 0:
                    continue
                self.H_expr.append(d2)
                h_row.append(j)
                h_col.append(k)
        self.G_index = np.array(g_index, dtype=np.int32)
        self.H_index_row = np.array(h_row, dtype=np.int32)
        self.H_index_col = np.array(h_col, dtype=np.int32)

    @property
    def n_G(self) -> int:
        return len(self.G_expr)

    @property
    def n_H(self) -> int:
        return len(self.H_expr)

    def free_symbols(self):
        return self.expr.free_symbols
