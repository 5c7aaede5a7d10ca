"""The caller side of the drop-in boundary for tests: SciPy trust-constr wired to a system's
callbacks exactly like the reference adapter (pockit/optimizer/scipy.py:13-29, 63-92) --
lower-triangle COO mirrored to the full Hessian with duplicates summed.  TEST INFRASTRUCTURE."""
import numpy as np
from scipy.optimize import Bounds, NonlinearConstraint, minimize
from scipy.sparse import coo_array


def _full(func, row, col, n):
    row, col = np.asarray(row), np.asarray(col)
    diag = np.nonzero(row == col)[0]

    def mat(*args):
        data = np.asarray(func(*args))
        half = coo_array((data, (row, col)), shape=(n, n))
        d = coo_array((data[diag], (row[diag], row[diag])), shape=(n, n))
        return half + half.T - d

    return mat


def solve(system, x0, options=None, objective=None):
    n, m = system.L, len(system.c_lb)
    jr, jc = system.jacobianstructure()
    cons = NonlinearConstraint(
        system.constraints, system.c_lb, system.c_ub,
        jac=lambda x: coo_array((system.jacobian(x), (jr, jc)), shape=(m, n)),
        hess=_full(system.hessian_c, *system.hessianstructure_c(), n),
    )
    return minimize(
        objective or system.objective, np.array(x0, dtype=np.float64), method="trust-constr", jac=system.gradient,
        hess=_full(system.hessian_o, *system.hessianstructure_o(), n), constraints=cons,
        bounds=Bounds(system.v_lb, system.v_ub), options=options or {},
    )
