"""Legendre-Gauss-Radau transcription (API of ``pockit.radau``:
``pockit/radau/phase.py:27-29``, ``pockit/radau/system.py:15-17``)."""
from .phase import Phase as _Phase
from .system import System as _System
from .guess import Variable, constant_guess, linear_guess  # noqa: F401


class Phase(_Phase):
    _scheme = "lgr"


class System(_System):
    _class_phase = Phase


__all__ = ["Phase", "System", "Variable", "constant_guess", "linear_guess"]
