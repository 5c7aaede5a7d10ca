"""Batched evaluation: ``B`` independent instances of one model per call.

The instances share the model, the mesh and therefore the sparsity patterns and the
compiled per-node programs; they differ in ``x`` (and multipliers) and, optionally,
in their FIXED boundary values (initial-condition sweeps: BASELINE.json's batched
planar_quadrotor configuration).  All buffers are instance-major ``[B][...]``.
"""
from __future__ import annotations

import numpy as np

from .phase import BcType

__all__ = ["BatchedSystem", "fixed_table", "fixed_index"]


def fixed_index(system, phase: int, kind: str, i: int = 0) -> int:
    """Column of the fixed-value table holding a FIXED boundary value.
    ``kind``: 'x0' (initial state ``i``), 'xf' (terminal state ``i``), 't0', 'tf'."""
    off = 0
    for p in system.p[:phase]:
        off += 2 * p.n_x + 2
    p = system.p[phase]
    info = {"x0": lambda: p.info_bc_0[i], "xf": lambda: p.info_bc_f[i], "t0": lambda: p.info_t_0,
            "tf": lambda: p.info_t_f}[kind]()
    if info.t != BcType.FIXED:
        raise ValueError(f"{kind}[{i}] of phase {phase} is not a FIXED boundary value")
    return off + {"x0": i, "xf": p.n_x + i, "t0": 2 * p.n_x, "tf": 2 * p.n_x + 1}[kind]


def fixed_table(system, batch: int) -> np.ndarray:
    """``[batch][n_fixed]`` table initialised with the model's own FIXED values."""
    row = []
    for p in system.p:
        for info in list(p.info_bc_0) + list(p.info_bc_f) + [p.info_t_0, p.info_t_f]:
            row.append(float(info.v) if info.t == BcType.FIXED else 0.0)
    return np.tile(np.array(row, dtype=np.float64), (batch, 1))


class BatchedSystem:
    def __init__(self, system, fixed: np.ndarray | None = None, batch: int | None = None, device: int | None = None):
        from .engine import Engine  # raises without the CUDA library / a device

        if fixed is None and batch is None:
            raise ValueError("give the fixed-value table or the batch size")
        self.system = system
        self.B = int(batch if fixed is None else len(fixed))
        self.fixed = fixed_table(system, self.B) if fixed is None else np.asarray(fixed, dtype=np.float64)
        self.engine = Engine(system.lowering, batch=self.B, fastmath=system._fastmath, device=device, fixed=self.fixed)
        self.L, self.m = system.L, len(system.c_lb)

    # structures are those of the single system
    def jacobianstructure(self):
        return self.system.jacobianstructure()

    def hessianstructure(self):
        return self.system.hessianstructure()

    def objective(self, X):
        return self.engine.objective(X)

    def gradient(self, X):
        return self.engine.gradient(X)

    def constraints(self, X):
        return self.engine.constraints(X)

    def jacobian(self, X):
        return self.engine.jacobian(X)

    def hessian(self, X, fct_c, fct_o):
        return self.engine.hessian(X, fct_c, fct_o)

    def close(self):
        self.engine.close()
