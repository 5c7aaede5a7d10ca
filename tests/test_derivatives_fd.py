"""CPU tier: the reference's own derivative tests (tests/test_labatto/test_derivative_lobatto.py,
tests/test_radau/test_derivative_radau.py) re-run against the planner + generated programs
(host-emulated): gradient, Jacobian and both Hessians of the reference's general test system
against central finite differences of the objective / constraints, with the reference's step sizes
and tolerances.  An independent pin next to the golden vectors: it does not involve the reference's
derivative code at all."""
import importlib

import numpy as np
import pytest

from hostemu import HostEmu
from pockit_b200 import plan as P
from pockit_b200 import problems


@pytest.fixture(scope="module", params=["lobatto", "radau"])
def general(request):
    S = problems.general(importlib.import_module(f"pockit_b200.{request.param}"))
    E = HostEmu(S)
    x = np.arange(S.L, dtype=np.float64) / 10 + 1  # the reference's test vector
    return S, E, x


def _dense(rows, cols, vals, shape):
    M = np.zeros(shape)
    np.add.at(M, (rows, cols), vals)
    return M


def test_gradient_and_jacobian_against_finite_differences(general):
    S, E, x = general
    n, m, eps = S.L, len(S.c_lb), 1e-6
    fd_g, fd_J = np.zeros(n), np.zeros((m, n))
    for i in range(n):
        xp, xm = x.copy(), x.copy()
        xp[i] += eps
        xm[i] -= eps
        fd_g[i] = (E.run(P.OBJ, xp)[0] - E.run(P.OBJ, xm)[0]) / (2 * eps)
        fd_J[:, i] = (E.run(P.CONS, xp) - E.run(P.CONS, xm)) / (2 * eps)
    assert np.allclose(E.run(P.GRAD, x), fd_g)
    assert np.allclose(_dense(*S.jacobianstructure(), E.run(P.JAC, x), (m, n)), fd_J)


def _fd_hessian(f, x, eps=2e-3):
    n = len(x)
    H = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            def at(di, dj):
                y = x.copy()
                y[i] += di * eps
                y[j] += dj * eps
                return f(y)
            H[i, j] = (at(1, 1) - at(1, -1) - at(-1, 1) + at(-1, -1)) / eps / eps / 4
    return H


def test_hessians_against_finite_differences(general):
    S, E, x = general
    n, m = S.L, len(S.c_lb)
    n_o = S.lowering.nnz_hess_o
    hr, hc = S.hessianstructure()
    # objective part: sigma = 1, lambda = 0
    sym = _dense(hr[:n_o], hc[:n_o], E.run(P.HESS, x, np.zeros(m), 1.0)[:n_o], (n, n))
    assert np.allclose(sym, _fd_hessian(lambda y: E.run(P.OBJ, y)[0], x), atol=1e-4, rtol=1e-4)
    # constraint part, one constraint at a time: lambda = e_c, sigma = 0
    for c in range(m):
        lam = np.zeros(m)
        lam[c] = 1.0
        sym = _dense(hr[n_o:], hc[n_o:], E.run(P.HESS, x, lam, 0.0)[n_o:], (n, n))
        fd = _fd_hessian(lambda y: E.run(P.CONS, y)[c], x)
        assert np.allclose(sym, fd, atol=1e-4, rtol=1e-4), f"constraint {c}"
