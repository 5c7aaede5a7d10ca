"""SciPy ``trust-constr`` adapter (caller side of the drop-in boundary, ``pockit/optimizer/scipy.py``):
same wiring -- lower-triangle COO Hessians mirrored to full matrices with duplicates summed,
Jacobian as a COO matrix -- in front of the CUDA engine, through the x-keyed cache."""
from __future__ import annotations

import numpy as np
from scipy.optimize import Bounds, NonlinearConstraint, minimize
from scipy.sparse import coo_array

from ._cache import CachedCallbacks
from ._common import pack_guess, unpack_solution

__all__ = ["solve"]


def _mirrored(func, row, col, n):
    """Full symmetric matrix from lower-triangle COO values (``scipy.py:13-29``)."""
    row, col = np.asarray(row), np.asarray(col)
    diag = np.flatnonzero(row == col)

    def matrix(*args):
        data = np.asarray(func(*args))
        half = coo_array((data, (row, col)), shape=(n, n))
        return half + half.T - coo_array((data[diag], (row[diag], row[diag])), shape=(n, n))

    return matrix


def solve(system, guess, optimizer_options=None, cache: bool = True):
    """Solve ``system`` with ``scipy.optimize.minimize(method="trust-constr")``.  ``guess`` and the
    return value follow the reference (a ``Variable``, or a list of them plus the static parameters);
    returns ``(result, scipy_result)``.  ``cache=False`` wires the system's callbacks directly."""
    x0, single, options = pack_guess(system, guess, optimizer_options)
    cb = CachedCallbacks(system) if cache else system
    n, m = system.L, len(system.c_lb)
    jr, jc = system.jacobianstructure()
    constraints = NonlinearConstraint(
        cb.constraints, system.c_lb, system.c_ub,
        jac=lambda x: coo_array((cb.jacobian(x), (jr, jc)), shape=(m, n)),
        hess=_mirrored(cb.hessian_c, *system.hessianstructure_c(), n),
    )
    res = minimize(
        cb.objective, x0, method="trust-constr", jac=cb.gradient,
        hess=_mirrored(cb.hessian_o, *system.hessianstructure_o(), n),
        constraints=constraints, bounds=Bounds(system.v_lb, system.v_ub), options=options,
    )
    if cache:
        res.cache_stats = dict(cb.stats)
    return unpack_solution(system, res.x, single), res
