"""x-keyed evaluation cache between a solver and the engine.

Ipopt and SciPy ask for the callbacks one at a time, but at the same point: ``f`` and ``g`` at every
trial point of a line search; ``grad f``, ``J`` and then ``H`` at every accepted one
(``optimizer/ipopt.py:41-53``, ``scipy.py:63-92``).  The reference recomputes the shared part of the
node graph for each; here

* ``x`` is copied to the device once per distinct point (later callbacks at that point run on the
  resident copy: ``pk_eval_set`` with ``x = NULL``), and
* callbacks that the solvers always request together are evaluated together, in one engine call
  with overlapped copies: ``{objective, constraints}`` and ``{gradient, jacobian}``.

Values are exactly the engine's.  The cache holds one point; a new ``x`` (compared element-wise,
NaNs compare unequal and therefore re-evaluate) drops everything.
"""
from __future__ import annotations

import numpy as np

from .. import plan as P

__all__ = ["CachedCallbacks"]

_GROUPS = {P.OBJ: (P.OBJ, P.CONS), P.CONS: (P.OBJ, P.CONS), P.GRAD: (P.GRAD, P.JAC), P.JAC: (P.GRAD, P.JAC)}


class CachedCallbacks:
    """``problem_obj`` for ``cyipopt.Problem`` / the callables ``scipy.optimize.minimize`` needs,
    in front of a :class:`pockit_b200.system.System`.  ``grouping=False`` evaluates one callback per
    request (still uploading ``x`` once per point)."""

    def __init__(self, system, grouping: bool = True):
        self.system = system
        self.engine = system.engine
        self.grouping = grouping
        self._x = None
        self._resident = False
        self._uploads = -1
        self._have: dict = {}
        self.stats = {"points": 0, "engine_calls": 0, "hits": 0}

    def __getattr__(self, name):  # bounds, layout, structures: the system's own
        return getattr(self.system, name)

    # ------------------------------------------------------------------
    def _at(self, x) -> None:
        x = np.asarray(x, dtype=np.float64)
        if self._x is not None and x.shape == self._x.shape and np.array_equal(x, self._x):
            return
        self._x = x.copy()
        self._resident = False
        self._have = {}
        self.stats["points"] += 1

    def _call(self, modes, fct_c=None, fct_o=None):
        # results land in the engine's page-locked buffers (device-to-host at full PCIe rate) and are
        # copied out once into the fresh arrays the solver keeps -- faster than copying 100 MB from the
        # device straight into pageable memory
        self.engine.reuse_outputs = True
        # The device copy of x is shared with every other entry point of the engine (System.objective,
        # System.evaluate, check_continuous, Engine.upload ...): residency is only trusted while the
        # engine's upload counter still reads what it read right after this cache's own upload.
        resident = self._resident and self.engine.x_uploads == self._uploads
        res = self.engine.evaluate(None if resident else self._x, fct_c, fct_o, modes=list(modes))
        self._resident = True
        self._uploads = self.engine.x_uploads
        self.stats["engine_calls"] += 1
        return res

    def _get(self, mode: int, x):
        self._at(x)
        if mode in self._have:
            self.stats["hits"] += 1
            return self._have[mode]
        group = [m for m in (_GROUPS[mode] if self.grouping else (mode,)) if m not in self._have]
        res = self._call(group)
        for m, v in res.items():
            self._have[m] = np.array(v, copy=True) if m != P.OBJ else v  # engine buffers may be reused
        return self._have[mode]

    # ------------------------------------------------------------------ callbacks (systembase.py:602-835)
    def objective(self, x):
        return self._get(P.OBJ, x)

    def gradient(self, x):
        return self._get(P.GRAD, x)

    def constraints(self, x):
        return self._get(P.CONS, x)

    def jacobian(self, x):
        return self._get(P.JAC, x)

    def hessian(self, x, fct_c, fct_o):
        self._at(x)
        return np.array(self._call([P.HESS], fct_c, fct_o)[P.HESS], copy=True)

    def _part(self, which: str, *args):
        # the engine evaluates the Hessian mode once and copies back only the requested part
        # (pk_download_range), at the resident point when this cache uploaded it
        self.engine.reuse_outputs = True
        shaped = bool(self.engine.compacted)
        resident = self._resident and self.engine.x_uploads == self._uploads and not shaped
        v = getattr(self.engine, which)(None if resident else self._x, *args)
        self._resident = True
        self._uploads = self.engine.x_uploads
        self.stats["engine_calls"] += 1
        return np.array(v, copy=True)

    def hessian_o(self, x):
        self._at(x)
        return self._part("hessian_o")

    def hessian_c(self, x, fct_c):
        self._at(x)
        return self._part("hessian_c", fct_c)
