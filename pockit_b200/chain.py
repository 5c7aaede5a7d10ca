"""Symbolic sparse chain rule.

The reference propagates derivatives through a small DAG of ``Node`` objects
whose "global" gradient / Hessian entries are *lists of NumPy arrays*
(``pockit/base/easyderiv.py:22-94``); every callback re-runs the value passes
``forward_gradient_v`` / ``forward_hessian_phase_v`` (:120-140, :231-304) on
freshly allocated arrays.  All of those passes only ever *multiply* per-node
quantities, so each list is fully described by

* an affine index pattern ``base + stride * k`` (``stride`` 0 = broadcast of a
  scalar column such as ``t_0`` or a static parameter, 1 = one column per node), and
* a product tree over a handful of *leaves* (function derivative at the node,
  ``dt``, the node's mesh fraction, boundary-function derivatives at ``s``).

This module builds exactly those descriptions once, at plan time.  The
traversal and the emission order reproduce ``composite_gradient_i`` (:97-108),
``composite_hessian_i_h`` (:143-158) and ``composite_hessian_phase_i_g``
(:161-193), so the resulting COO patterns are the reference's, duplicates and
order included, while the values become straight-line device code.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence, Union

__all__ = ["Leaf", "Prod", "Term", "ONE", "GEntry", "HEntry", "Sym", "compose", "precedes"]


@dataclass(frozen=True)
class Leaf:
    """A primitive factor.  ``kind`` selects the meaning of ``key``:

    ``F`` / ``G`` / ``H``   value / first / second derivative number ``key[2]`` of
                            function ``key[1]`` in family ``key[0]``, at the node
    ``bG`` / ``bH``         derivative of a boundary function (of ``s`` only)
    ``dt``                  ``t_f - t_0``
    ``tm`` / ``om``         mesh fraction of the node / one minus it
    ``sG`` / ``sH``         derivative of a system-level function (objective, system
                            constraint) w.r.t. integrals / static parameters
    """

    kind: str
    key: tuple = ()


@dataclass(frozen=True)
class Prod:
    left: "Tree"
    right: "Tree"


Tree = Union[Leaf, Prod, None]  # None is the multiplicative identity 1.0


@dataclass(frozen=True)
class Term:
    """``coef * tree`` where ``coef`` is ±1 or ±2 (exact in binary floating
    point, so its position in the product does not matter)."""

    tree: Tree = None
    coef: float = 1.0

    def __mul__(self, other: "Term") -> "Term":
        if self.tree is None:
            tree = other.tree
        elif other.tree is None:
            tree = self.tree
        else:
            tree = Prod(self.tree, other.tree)
        return Term(tree, self.coef * other.coef)

    def scaled(self, c: float) -> "Term":
        return Term(self.tree, self.coef * c)

    def leaves(self):
        out = []

        def walk(t):
            if t is None:
                return
            if isinstance(t, Leaf):
                out.append(t)
            else:
                walk(t.left)
                walk(t.right)

        walk(self.tree)
        return out


ONE = Term()


def leaf(kind: str, *key) -> Term:
    return Term(Leaf(kind, tuple(key)))


@dataclass
class GEntry:
    base: int
    stride: int
    val: Term


@dataclass
class HEntry:
    row_base: int
    row_stride: int
    col_base: int
    col_stride: int
    val: Term


@dataclass
class Sym:
    """Symbolic twin of an easyderiv ``Node``: only the global lists survive."""

    G: list[GEntry] = field(default_factory=list)
    H: list[HEntry] = field(default_factory=list)


def precedes(a: int, b: int) -> bool:
    """Column order with static parameters (negative, phase-local) after all
    phase variables -- ``easyderiv.py:8-19``."""
    if a < 0:
        return b < 0 and a < b
    return b < 0 or a < b


def compose(
    args: Sequence[Sym],
    g_index: Sequence[int],
    g_val: Sequence[Term],
    h_row: Sequence[int] = (),
    h_col: Sequence[int] = (),
    h_val: Sequence[Term] = (),
) -> Sym:
    """Lists of ``f(args)`` given the local derivative leaves of ``f``."""
    out = Sym()
    for jj, j in enumerate(g_index):
        for a in args[j].G:
            out.G.append(GEntry(a.base, a.stride, a.val * g_val[jj]))
    # first-derivative x argument-Hessian
    for jj, j in enumerate(g_index):
        for a in args[j].H:
            out.H.append(HEntry(a.row_base, a.row_stride, a.col_base, a.col_stride, a.val * g_val[jj]))
    # second-derivative x argument-gradient x argument-gradient
    for m, (r, c) in enumerate(zip(h_row, h_col)):
        for a in args[r].G:
            for b in args[c].G:
                val = a.val * b.val * h_val[m]
                if r == c:
                    if precedes(a.base, b.base):
                        continue
                    out.H.append(HEntry(a.base, a.stride, b.base, b.stride, val))
                else:
                    if a.base == b.base:
                        val = val.scaled(2.0)
                    if precedes(a.base, b.base):
                        out.H.append(HEntry(b.base, b.stride, a.base, a.stride, val))
                    else:
                        out.H.append(HEntry(a.base, a.stride, b.base, b.stride, val))
    return out
