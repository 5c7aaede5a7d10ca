"""Solver adapters: the callers of the hot path (``pockit/optimizer/ipopt.py``, ``scipy.py``),
re-stated for the B200 engine with an x-keyed evaluation cache between the solver
and the engine."""
from . import ipopt, scipy  # noqa: F401
from ._cache import CachedCallbacks  # noqa: F401
