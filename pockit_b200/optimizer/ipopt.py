"""Ipopt adapter (caller side of the drop-in boundary, ``pockit/optimizer/ipopt.py:11-61``):
``cyipopt.Problem`` over the system's callbacks, through the x-keyed cache.  ``cyipopt`` is imported
on use (it is not part of this image; the adapter raises ``ImportError`` without it)."""
from __future__ import annotations

from ._cache import CachedCallbacks
from ._common import pack_guess, unpack_solution

__all__ = ["solve"]


def solve(system, guess, optimizer_options=None, cache: bool = True):
    """Solve ``system`` with Ipopt; options are passed through unchanged.  Returns ``(result, info)``."""
    try:
        import cyipopt
    except ImportError as exc:  # pragma: no cover - depends on the environment
        raise ImportError("pockit_b200.optimizer.ipopt needs cyipopt (and libipopt), which is not installed") from exc
    x0, single, options = pack_guess(system, guess, optimizer_options)
    cb = CachedCallbacks(system) if cache else system
    problem = cyipopt.Problem(n=int(system.L), m=len(system.c_lb), problem_obj=cb, lb=system.v_lb, ub=system.v_ub,
                              cl=system.c_lb, cu=system.c_ub)
    for k, v in options.items():
        problem.add_option(k, v)
    x, info = problem.solve(x0)
    if cache:
        info["cache_stats"] = dict(cb.stats)
    return unpack_solution(system, x, single), info
