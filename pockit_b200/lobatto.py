"""Legendre-Gauss-Lobatto transcription (API of ``pockit.lobatto``:
``pockit/lobatto/phase.py:28-30``, ``pockit/lobatto/system.py:18-20``)."""
from .phase import Phase as _Phase
from .system import System as _System
from .guess import Variable, constant_guess, linear_guess  # noqa: F401


class Phase(_Phase):
    _scheme = "lgl"


class System(_System):
    _class_phase = Phase


__all__ = ["Phase", "System", "Variable", "constant_guess", "linear_guess"]
